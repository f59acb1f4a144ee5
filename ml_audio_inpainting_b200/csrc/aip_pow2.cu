// Any power-of-two n_fft from 64 to 2048 other than the fused 512 path -- the reference's DEFAULTS are among them
// (utils.extract_spectrogram / extract_mel_spectrogram: n_fft 2048, hop 512, utils.py:192-193, :236-238; the tests use 256, 1024, 2048).
//
//   stft_pow2_fwd_kernel      tile = FT consecutive frames of one clip; a WARP transforms a "super-frame" of 1024 complex points
//                             (1024 / M frames of M = n_fft / 2 packed points) by Stockham autosort passes of radix 16, 16 and
//                             r3 = M / 16 / 16 (or M / 16) whose butterflies live in registers -- the packed fft16x2 codelet of the
//                             512 kernels, two butterflies per lane -- then the real-input split pass in place; once the tile's
//                             frames sit in shared memory in natural bin order, the CTA writes them out TRANSPOSED: FT consecutive
//                             frames of one bin are contiguous in the [F, T] output, so every store covers full 32-byte sectors.
//   istft_pow2_kernel         the inverse of the same steps: transposed (coalesced) load of FT frames' bins, inverse split pass,
//                             the same forward passes on (im, re)-swapped data, window and 1 / M, and the overlap-add of the
//                             tile's frames straight from shared memory (very small hops: frames to the workspace and
//                             istft_generic_ola_kernel, aip_inv.cu).
//
// The one-frame-per-CTA radix-2 kernels they replace (stft_generic_fwd_kernel, istft_generic_frames_kernel: 0.24 TB/s on
// 512 x 10 s clips, isolated 8-byte stores, sincospif per butterfly) remain for n_fft 32 and 4096 and behind AIP_POW2=0.
//
// Shared-memory layout of one frame: M + 1 complex values at index a + (a >> 4) (one pad slot per 16: the radix-16 passes store
// 16 consecutive values per lane and read with stride M / 16, both conflict-free this way), pitch PB >= M + M / 16 + 1 float2, which
// is chosen so that the transposed access (lanes along the frames) is conflict-free as well (pow2_geom).
#include "aip_device.cuh"
#include "aip_host.h"

namespace aip {

constexpr int kP2LoadUnroll = 8;      // complex rows a thread of the tiled inverse keeps in flight (16 / 32 measured: 2048 slower, 1024 3 % faster)
constexpr int kP2Threads = 256;
constexpr int kP2Warps = kP2Threads / 32;
constexpr int kP2Points = 1024;                 // complex points per super-frame (32 lanes x 2 packed butterflies x 16)

struct Pow2Geom {
  int N, M;            // n_fft, packed complex length N / 2
  int logJ;            // J = M / 16 radix-16 butterflies per frame and pass
  int fps;             // frames per super-frame = 1024 / M
  int FT, logFT;       // frames per tile = 8 warps * fps
  int PB;              // float2 pitch of a frame buffer
  int two16;           // second radix-16 pass (M >= 256)
  int r3, logNs3;      // radix of the last pass (1: none) and log2 of its sub-transform length Ns = M / r3
};

__host__ __device__ constexpr int p2_ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
// The kernels are instantiated per n_fft with this as a compile-time constant: every stride, trip count and shift below folds.
__host__ __device__ constexpr Pow2Geom pow2_geom(int n_fft) {
  Pow2Geom g{};
  g.N = n_fft; g.M = n_fft / 2;
  g.logJ = p2_ilog2(g.M) - 4;
  g.fps = kP2Points / g.M;
  g.FT = kP2Warps * g.fps; g.logFT = p2_ilog2(g.FT);
  // transposed access: a half-warp reads FT (or 16) frames x 16 / FT bins as 8-byte words -> PB == 16 / min(FT, 16) (mod 16)
  g.PB = g.M + g.M / 16 + (g.FT == 8 ? 2 : 1);
  g.two16 = g.M >= 256;
  g.r3 = g.two16 ? g.M / 256 : g.M / 16;
  g.logNs3 = g.r3 > 1 ? p2_ilog2(g.M / g.r3) : 0;
  return g;
}

__device__ __forceinline__ int p2_pad(int a) { return a + (a >> 4); }

// Shared-memory tables of a CTA (floats): tw [2 M] = W_N^e, e < M (split pass) | win [N] | tw2 [512] = W_256^(m k) at [16 m + k]
// (second radix-16 pass: lanes k = 0..15 read consecutive entries) | twl [512] = W_M^j, j < Ns (last pass: lane j reads entry j
// and raises it to the powers 2 .. R-1 itself -- indexing one big table with r j is a 2r-way bank conflict) | frame buffers
constexpr int kP2TabFloats = 2 * 256 + 2 * 256;
struct P2Smem {
  float2* tw; float* win; float2* tw2; float2* twl; float2* bufs;
};
__device__ __forceinline__ P2Smem p2_smem(const Pow2Geom& g, float* smem) {
  P2Smem m;
  m.tw = reinterpret_cast<float2*>(smem);
  m.win = smem + 2 * g.M;
  m.tw2 = reinterpret_cast<float2*>(smem + 2 * g.M + g.N);
  m.twl = m.tw2 + 256;
  m.bufs = m.twl + 256;
  return m;
}
__device__ __forceinline__ float2 p2_cis(int num, int den) {       // exp(-2 pi j num / den)
  float sn, cs;
  sincospif(2.0f * (float)num / (float)den, &sn, &cs);
  return make_float2(cs, -sn);
}
__device__ __forceinline__ void p2_fill_tables(const Pow2Geom& g, const float* window, const P2Smem& m, float win_scale) {
  for (int e = threadIdx.x; e < g.M; e += blockDim.x) m.tw[e] = p2_cis(e, g.N);
  for (int n = threadIdx.x; n < g.N; n += blockDim.x) m.win[n] = window[n] * win_scale;
  for (int e = threadIdx.x; e < 256; e += blockDim.x) {
    m.tw2[e] = p2_cis(((e >> 4) * (e & 15)) & 255, 256);
    m.twl[e] = p2_cis(e, g.M);
  }
}

// ---- radix-16 passes: lane holds butterflies (slots) `lane` and `lane + 32` of the super-frame, packed in .x / .y ----------------
struct P2Slots {
  int fA, jA, fB, jB;      // frame-in-super-frame and butterfly index of the two slots
};
__device__ __forceinline__ P2Slots p2_slots(const Pow2Geom& g, int lane) {
  P2Slots s;
  const int J1 = (1 << g.logJ) - 1;
  s.fA = lane >> g.logJ; s.jA = lane & J1;
  s.fB = (lane + 32) >> g.logJ; s.jB = (lane + 32) & J1;
  return s;
}

// first pass (sub-transform length 1: no twiddles).  load.a(idx) / load.b(idx) -> packed point idx of the frame of slot A / B.
template <class Load>
__device__ __forceinline__ void p2_pass_first(const Pow2Geom& g, const P2Slots& s, float2* sf, const Load& load) {
  float2 r[16], i[16];
  const int J = 1 << g.logJ;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const float2 a = load.a(s.jA + J * m), b = load.b(s.jB + J * m);
    r[m] = make_float2(a.x, b.x);
    i[m] = make_float2(a.y, b.y);
  }
  fft16x2(r, i);
  __syncwarp();                                   // the inverse reads the same buffers: every load before any store
  float2* oA = sf + s.fA * g.PB + 17 * s.jA;      // p2_pad(16 j + k1) = 17 j + k1
  float2* oB = sf + s.fB * g.PB + 17 * s.jB;
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    oA[k1] = make_float2(r[perm16(k1)].x, i[perm16(k1)].x);
    oB[k1] = make_float2(r[perm16(k1)].y, i[perm16(k1)].y);
  }
  __syncwarp();
}

// second radix-16 pass (sub-transform length 16 -> 256), in place
__device__ __forceinline__ void p2_pass_second(const Pow2Geom& g, const P2Slots& s, float2* sf, const float2* tw2) {
  float2 r[16], i[16];
  const int J = 1 << g.logJ;                      // a multiple of 16 here
  const float2* iA = sf + s.fA * g.PB + s.jA + (s.jA >> 4);
  const float2* iB = sf + s.fB * g.PB + s.jB + (s.jB >> 4);
  const int stride = J + (J >> 4);
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const float2 a = iA[stride * m], b = iB[stride * m];
    r[m] = make_float2(a.x, b.x);
    i[m] = make_float2(a.y, b.y);
  }
  const int k = s.jA & 15;                        // == jB & 15: both slots share the twiddles W_256^(m k)
#pragma unroll
  for (int m = 1; m < 16; ++m) {
    const float2 w = tw2[16 * m + k];
    cmulx2(r[m], i[m], w.x, w.y);
  }
  fft16x2(r, i);
  __syncwarp();
  float2* oA = sf + s.fA * g.PB + p2_pad(((s.jA >> 4) << 8) + k);
  float2* oB = sf + s.fB * g.PB + p2_pad(((s.jB >> 4) << 8) + k);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {               // p2_pad(j0 + 16 k1) = p2_pad(j0) + 17 k1 (j0 % 16 + 16 k1 never carries twice)
    oA[17 * k1] = make_float2(r[perm16(k1)].x, i[perm16(k1)].x);
    oB[17 * k1] = make_float2(r[perm16(k1)].y, i[perm16(k1)].y);
  }
  __syncwarp();
}

// ---- last pass: radix R3 in {2, 4, 8}, sub-transform length Ns = M / R3 -> M; in place by construction --------------------------
template <int R>
__device__ __forceinline__ void p2_dft_small(float (&xr)[R], float (&xi)[R]);
template <>
__device__ __forceinline__ void p2_dft_small<2>(float (&xr)[2], float (&xi)[2]) {
  const float ar = xr[0], ai = xi[0];
  xr[0] = ar + xr[1]; xi[0] = ai + xi[1];
  xr[1] = ar - xr[1]; xi[1] = ai - xi[1];
}
template <>
__device__ __forceinline__ void p2_dft_small<4>(float (&xr)[4], float (&xi)[4]) {
  const float s0r = xr[0] + xr[2], s0i = xi[0] + xi[2], s1r = xr[0] - xr[2], s1i = xi[0] - xi[2];
  const float s2r = xr[1] + xr[3], s2i = xi[1] + xi[3], s3r = xr[1] - xr[3], s3i = xi[1] - xi[3];
  xr[0] = s0r + s2r; xi[0] = s0i + s2i;
  xr[2] = s0r - s2r; xi[2] = s0i - s2i;
  xr[1] = s1r + s3i; xi[1] = s1i - s3r;           // s1 - j s3
  xr[3] = s1r - s3i; xi[3] = s1i + s3r;           // s1 + j s3
}
template <>
__device__ __forceinline__ void p2_dft_small<8>(float (&xr)[8], float (&xi)[8]) {
  float er[4] = {xr[0], xr[2], xr[4], xr[6]}, ei[4] = {xi[0], xi[2], xi[4], xi[6]};
  float qr[4] = {xr[1], xr[3], xr[5], xr[7]}, qi[4] = {xi[1], xi[3], xi[5], xi[7]};
  p2_dft_small<4>(er, ei);
  p2_dft_small<4>(qr, qi);
  // W8^k q[k]: 1, (1 - j) / sqrt 2, -j, (-1 - j) / sqrt 2
  const float t1r = (qr[1] + qi[1]) * kR2, t1i = (qi[1] - qr[1]) * kR2;
  const float t2r = qi[2], t2i = -qr[2];
  const float t3r = (qi[3] - qr[3]) * kR2, t3i = -(qr[3] + qi[3]) * kR2;
  xr[0] = er[0] + qr[0]; xi[0] = ei[0] + qi[0]; xr[4] = er[0] - qr[0]; xi[4] = ei[0] - qi[0];
  xr[1] = er[1] + t1r;   xi[1] = ei[1] + t1i;   xr[5] = er[1] - t1r;   xi[5] = ei[1] - t1i;
  xr[2] = er[2] + t2r;   xi[2] = ei[2] + t2i;   xr[6] = er[2] - t2r;   xi[6] = ei[2] - t2i;
  xr[3] = er[3] + t3r;   xi[3] = ei[3] + t3i;   xr[7] = er[3] - t3r;   xi[7] = ei[3] - t3i;
}

template <int R>
__device__ __forceinline__ void p2_pass_last(const Pow2Geom& g, float2* sf, const float2* twl, int lane) {
  const int Ns = 1 << g.logNs3;
  const int stride = p2_pad(Ns);                  // Ns is a multiple of 16: p2_pad(j + r Ns) = p2_pad(j) + r p2_pad(Ns)
#pragma unroll
  for (int q = lane; q < kP2Points / R; q += 32) {
    const int f = q >> g.logNs3, j = q & (Ns - 1);
    float2* p = sf + f * g.PB + p2_pad(j);
    float xr[R], xi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float2 v = p[stride * r];
      xr[r] = v.x; xi[r] = v.y;
    }
    const float2 w1 = twl[j];                                     // W_M^j; W_M^(r j) by repeated multiplication
    float wr = w1.x, wi = w1.y;
#pragma unroll
    for (int r = 1; r < R; ++r) {
      cmul(xr[r], xi[r], wr, wi);
      if (r + 1 < R) cmul(wr, wi, w1.x, w1.y);
    }
    p2_dft_small<R>(xr, xi);
#pragma unroll
    for (int r = 0; r < R; ++r) p[stride * r] = make_float2(xr[r], xi[r]);
  }
  __syncwarp();
}

// passes after the first one
__device__ __forceinline__ void p2_passes_rest(const Pow2Geom& g, const P2Slots& s, float2* sf, const P2Smem& m, int lane) {
  if (g.two16) p2_pass_second(g, s, sf, m.tw2);
  if (g.r3 == 2) p2_pass_last<2>(g, sf, m.twl, lane);
  else if (g.r3 == 4) p2_pass_last<4>(g, sf, m.twl, lane);
  else if (g.r3 == 8) p2_pass_last<8>(g, sf, m.twl, lane);
}

// ---- forward -----------------------------------------------------------------------------------------------------------------
struct Pow2FwdParams {
  FwdParams P;
  int F;               // bins = M + 1
  int vec;             // float2 loads of the waveform are legal (even hop / pad / pitch, 8-byte aligned base)
  int span;            // > 0: floats of the tile's waveform span (FT - 1) hop + N, staged ONCE in shared memory by a bulk copy
                       //      that runs while the previous tile is written out; 0: every warp loads its frames from global
};

// Waveform loaders of the first pass: the lane's two frames start at clip samples gA and gB.  P2WaveClean: both frames lie inside
// [0, L) and outside the gap and 8-byte loads are legal -- 32 independent loads, nothing else; P2WaveChecked: any frame (zero
// padding, gap, ragged end, odd hop), still branch-free so that the loads stay independent: clamped address, value selected.
struct P2WaveClean {
  const float* __restrict__ src;
  const float* win;
  int gA, gB;
  __device__ __forceinline__ float2 get(int g0, int idx) const {
    const float2 v = __ldg(reinterpret_cast<const float2*>(src + g0 + 2 * idx));
    const float2 t = *reinterpret_cast<const float2*>(win + 2 * idx);
    return make_float2(v.x * t.x, v.y * t.y);
  }
  __device__ __forceinline__ float2 a(int idx) const { return get(gA, idx); }
  __device__ __forceinline__ float2 b(int idx) const { return get(gB, idx); }
};
struct P2WaveChecked {
  const float* __restrict__ src;
  const float* win;
  int gA, gB, L, gs, ge, reflect;
  __device__ __forceinline__ float one(int g) const {
    if (reflect && (g < 0 || g >= L)) g = g < 0 ? -g : 2 * (L - 1) - g;      // np.pad(mode="reflect") of the gapped clip
    const bool ok = g >= 0 && g < L && !(g >= gs && g < ge);
    const float v = __ldg(src + (ok ? g : 0));
    return ok ? v : 0.0f;
  }
  __device__ __forceinline__ float2 get(int g0, int idx) const {
    const float2 t = *reinterpret_cast<const float2*>(win + 2 * idx);
    return make_float2(one(g0 + 2 * idx) * t.x, one(g0 + 2 * idx + 1) * t.y);
  }
  __device__ __forceinline__ float2 a(int idx) const { return get(gA, idx); }
  __device__ __forceinline__ float2 b(int idx) const { return get(gB, idx); }
};

// ... and from the staged span (already zero-padded and gap-zeroed by the bulk copy + fwd_fixup): pa / pb = the frames' first sample
struct P2WaveStaged {
  const float* pa;
  const float* pb;
  const float* win;
  __device__ __forceinline__ float2 get(const float* p, int idx) const {
    const float2 v = *reinterpret_cast<const float2*>(p + 2 * idx);
    const float2 t = *reinterpret_cast<const float2*>(win + 2 * idx);
    return make_float2(v.x * t.x, v.y * t.y);
  }
  __device__ __forceinline__ float2 a(int idx) const { return get(pa, idx); }
  __device__ __forceinline__ float2 b(int idx) const { return get(pb, idx); }
};

// the span of tile (b, t0) as a bulk-copy plan (aip_tiles.cuh: the in-range, 16-byte part moves by ONE cp.async.bulk, the zero
// padding, a ragged end and the gap are patched by fwd_fixup once it has landed)
__device__ __forceinline__ FwdTilePlan p2_span_plan(const FwdParams& P, int b, int t0, int span, int gs, int ge) {
  FwdTilePlan q;
  q.len = span;
  q.g0 = t0 * P.hop - P.pad;
  q.src = P.wave + (long long)b * P.wave_pitch;
  q.v_lo = q.g0 < 0 ? -q.g0 : 0;
  const int hi = P.L - q.g0;
  q.v_hi = hi < q.len ? (hi > 0 ? hi : 0) : q.len;
  if (q.v_lo > q.v_hi) q.v_lo = q.v_hi;
  q.n_bulk = (q.v_hi - q.v_lo) & ~3;
  q.gs = gs; q.ge = ge;
  q.L = P.L; q.reflect = P.reflect;
  return q;
}
__device__ __forceinline__ void p2_span_issue(const FwdTilePlan& q, float* span, uint64_t* bar) {
  if (q.n_bulk > 0) {
    const uint32_t bytes = (uint32_t)q.n_bulk * 4u;
    mbar_expect_tx(bar, bytes);
    tma_load_1d(span + q.v_lo, q.src + q.g0 + q.v_lo, bytes, bar);
  } else {
    mbar_arrive(bar);
  }
}

// real-input split pass of one super-frame, in place: Z (M packed points) -> X[0..M]
__device__ __forceinline__ void p2_split_fwd(const Pow2Geom& g, float2* sf, const float2* tw, int lane) {
  const int H = g.M >> 1, logH = g.logJ + 3;
#pragma unroll 4
  for (int q = lane; q < kP2Points / 2; q += 32) {
    const int f = q >> logH, k = q & (H - 1);
    float2* z = sf + f * g.PB;
    if (k == 0) {
      const float2 z0 = z[0], zh = z[p2_pad(H)];
      z[0] = make_float2(z0.x + z0.y, 0.0f);
      z[p2_pad(g.M)] = make_float2(z0.x - z0.y, 0.0f);
      z[p2_pad(H)] = make_float2(zh.x, -zh.y);
    } else {
      const float2 a = z[p2_pad(k)], b = z[p2_pad(g.M - k)], w = tw[k];
      const float er = 0.5f * (a.x + b.x), ei = 0.5f * (a.y - b.y);
      const float orr = 0.5f * (a.y + b.y), oi = -0.5f * (a.x - b.x);
      const float tr = w.x * orr - w.y * oi, ti = w.x * oi + w.y * orr;
      z[p2_pad(k)] = make_float2(er + tr, ei + ti);
      z[p2_pad(g.M - k)] = make_float2(er - tr, -(ei - ti));
    }
  }
}

template <int kN>
__global__ void __launch_bounds__(kP2Threads, 2) stft_pow2_fwd_kernel(const Pow2FwdParams G) {
  extern __shared__ __align__(16) float smem[];
  constexpr Pow2Geom g = pow2_geom(kN);
  const FwdParams& P = G.P;
  const P2Smem sm = p2_smem(g, smem);
  const float2* tw = sm.tw;
  const float* win = sm.win;
  float2* bufs = sm.bufs;
  __shared__ __align__(8) uint64_t span_bar;
  float* span = reinterpret_cast<float*>(bufs + g.FT * g.PB);
  p2_fill_tables(g, P.window, sm, 1.0f);
  if (threadIdx.x == 0) mbar_init(&span_bar, 1);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const P2Slots s = p2_slots(g, lane);
  float2* sf = bufs + warp * g.fps * g.PB;
  uint32_t phase = 0;
  if (G.span > 0 && threadIdx.x == 0 && (int)blockIdx.x < P.n_tiles) {
    const int b = blockIdx.x / P.tiles_per_clip, t0 = (blockIdx.x - b * P.tiles_per_clip) * g.FT;
    p2_span_issue(p2_span_plan(P, b, t0, G.span, 0, 0), span, &span_bar);
  }
  for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const int b = tile / P.tiles_per_clip, t0 = (tile - b * P.tiles_per_clip) * g.FT;
    int gs = 0, ge = 0;
    if (P.gap_samples) { gs = P.gap_samples[2 * b]; ge = P.gap_samples[2 * b + 1]; }
    const float* src = P.wave + (long long)b * P.wave_pitch;
    {
      int tA = t0 + warp * g.fps + s.fA, tB = t0 + warp * g.fps + s.fB;
      if (tA > P.T_out - 1) tA = P.T_out - 1;                     // frames past the end replay the last one; never stored
      if (tB > P.T_out - 1) tB = P.T_out - 1;
      const int gA = tA * P.hop - P.pad, gB = tB * P.hop - P.pad;
      const bool clean = gA >= 0 && gA + g.N <= P.L && (ge <= gs || ge <= gA || gs >= gA + g.N) &&
                         gB >= 0 && gB + g.N <= P.L && (ge <= gs || ge <= gB || gs >= gB + g.N);
      if (G.span > 0) {
        const FwdTilePlan q = p2_span_plan(P, b, t0, G.span, gs, ge);
        mbar_wait(&span_bar, phase);
        phase ^= 1;
        if (fwd_needs_fixup(q)) {
          fwd_fixup(q, threadIdx.x, span);
          __syncthreads();
        }
        const P2WaveStaged load{span + (tA - t0) * P.hop, span + (tB - t0) * P.hop, win};
        p2_pass_first(g, s, sf, load);
      } else if (G.vec && __all_sync(0xffffffffu, clean)) {
        const P2WaveClean load{src, win, gA, gB};
        p2_pass_first(g, s, sf, load);
      } else {
        const P2WaveChecked load{src, win, gA, gB, P.L, gs, ge, P.reflect};
        p2_pass_first(g, s, sf, load);
      }
    }
    p2_passes_rest(g, s, sf, sm, lane);
    p2_split_fwd(g, sf, tw, lane);
    if (G.span > 0) fence_proxy_async();             // this thread's reads of the span come before the next copy into it
    __syncthreads();
    if (G.span > 0 && threadIdx.x == 0) {            // the next tile's span arrives while this one is written out
      const int nt = tile + gridDim.x;
      if (nt < P.n_tiles) {
        const int nb = nt / P.tiles_per_clip, nt0 = (nt - nb * P.tiles_per_clip) * g.FT;
        p2_span_issue(p2_span_plan(P, nb, nt0, G.span, 0, 0), span, &span_bar);
      }
    }
    // transposed store: thread = (frame f, bins k0 + i * 256 / FT); the epilogue switches are CTA-uniform
    {
      const int f = threadIdx.x & (g.FT - 1), k0 = threadIdx.x >> g.logFT, kstep = kP2Threads >> g.logFT;
      const int t = t0 + f;
      if (t < P.T_out) {
        bool zero = false;
        if (P.zero_frames) zero = (t >= P.zero_frames[2 * b] && t < P.zero_frames[2 * b + 1]);
        float maskv = 0.0f;
        if (P.mask) {
          bool in = false;
          if (P.mask_frames) in = (t >= P.mask_frames[2 * b] && t < P.mask_frames[2 * b + 1]);
          maskv = (in == (P.mask_in_gap_is_one != 0)) ? 1.0f : 0.0f;
        }
        const float2* z = bufs + f * g.PB;
        const long long step = (long long)kstep * P.T_out;
        long long idx = (long long)b * G.F * P.T_out + t + (long long)k0 * P.T_out;
        if (P.spec && !P.phase && !P.mask && P.mag_kind == MAG_NONE && !P.zero_frames) {       // complex output only
          float2* o = P.spec + idx;
#pragma unroll 8
          for (int k = k0; k <= g.M; k += kstep, o += step) *o = z[p2_pad(k)];
        } else if (!P.spec && !P.phase && !P.mask && !P.zero_frames && P.mag_kind != MAG_NONE) {   // one magnitude flavour only
          float* o = P.mag + idx;
          const int mk = P.mag_kind;
          const float eps = P.eps, power = P.power;
#pragma unroll 4
          for (int k = k0; k <= g.M; k += kstep, o += step) {
            const float2 v = z[p2_pad(k)];
            *o = mag_value(mk, v.x, v.y, eps, power);
          }
        } else {
#pragma unroll 2
          for (int k = k0; k <= g.M; k += kstep, idx += step) {
            float2 v = z[p2_pad(k)];
            if (zero) v = make_float2(0.0f, 0.0f);
            if (P.spec) P.spec[idx] = v;
            if (P.phase) P.phase[idx] = fast_atan2(v.y, v.x);
            if (P.mask) P.mask[idx] = maskv;
            if (P.mag_kind != MAG_NONE) P.mag[idx] = mag_value(P.mag_kind, v.x, v.y, P.eps, P.power);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---- inverse: frames to the workspace --------------------------------------------------------------------------------------------
struct Pow2InvParams {
  InvParams P;
  int F;
  float* frames;       // workspace [B, T, N]
  int ola_k, log_hop;  // fused overlap-add, fast form: hop = N / ola_k is a power of two <= 512 (ola_k in 1..8); 0: general gather
};

struct P2SmemLoad {
  const float2* fa;
  const float2* fb;    // the buffers of the lane's two frames
  __device__ __forceinline__ float2 a(int idx) const { return fa[p2_pad(idx)]; }
  __device__ __forceinline__ float2 b(int idx) const { return fb[p2_pad(idx)]; }
};

// inverse split pass, in place: X[0..M] -> Z (M packed points), stored (im, re)-SWAPPED so that the forward passes compute the
// inverse transform (swap(IDFT(Z)) * M = DFT(swap(Z))).  scipy.fft.irfft ignores imag(DC) and imag(Nyquist); so does this.
__device__ __forceinline__ void p2_split_inv(const Pow2Geom& g, float2* sf, const float2* tw, int lane) {
  const int H = g.M >> 1, logH = g.logJ + 3;
#pragma unroll 4
  for (int q = lane; q < kP2Points / 2; q += 32) {
    const int f = q >> logH, k = q & (H - 1);
    float2* z = sf + f * g.PB;
    if (k == 0) {
      const float x0 = z[0].x, xm = z[p2_pad(g.M)].x;
      const float2 xh = z[p2_pad(H)];
      z[0] = make_float2(0.5f * (x0 - xm), 0.5f * (x0 + xm));            // swapped: (im, re)
      z[p2_pad(H)] = make_float2(-xh.y, xh.x);                           // conj(X[M/2]) swapped
    } else {
      const float2 a = z[p2_pad(k)], b = z[p2_pad(g.M - k)], w = tw[k];
      const float er = 0.5f * (a.x + b.x), ei = 0.5f * (a.y - b.y);
      const float dr = 0.5f * (a.x - b.x), di = 0.5f * (a.y + b.y);      // (X[k] - conj(X[M-k])) / 2
      const float orr = w.x * dr + w.y * di, oi = w.x * di - w.y * dr;   // times conj(W_N^k)
      // Z[k] = E + j O ; Z[M-k] = conj(E) + j conj(O)
      z[p2_pad(k)] = make_float2(ei + orr, er - oi);
      z[p2_pad(g.M - k)] = make_float2(orr - ei, er + oi);
    }
  }
  __syncwarp();
}

// Overlap-add gather for hop = N / kK, a power of two <= 512: a thread's pairs u = 2 tid + 512 it all sit at the same offset
// r inside their hop, so the kK window taps and buffer offsets it needs are the same for all of them (w / off) and the frames
// covering a pair are q - kK + 1 .. q with q = u / hop = q0 + it (512 / hop): no division, no address arithmetic in the loop.
template <int kK>
__device__ __forceinline__ void p2_gather_pow2hop(const InvParams& P, const float2* bufs, const float* win, int PB, int FT, int log_hop,
                                                  int t0, int nv, int span, float* orow) {
  // (set up per tile, not once per kernel: 3 kK registers that would otherwise stay live across the transform passes)
  float2 w[kK];
  int off[kK];
  {
    const int r = (2 * threadIdx.x) & (P.hop - 1);
#pragma unroll
    for (int j = 0; j < kK; ++j) {
      const int n = j * P.hop + r;
      w[j] = *reinterpret_cast<const float2*>(win + n);
      off[j] = p2_pad(n >> 1);
    }
  }
  const int p0 = t0 * P.hop;
  const int dq = 512 >> log_hop;
  for (int u = 2 * threadIdx.x, q = (2 * threadIdx.x) >> log_hop; u < span; u += 2 * kP2Threads, q += dq) {
    const int sidx = p0 + u - P.pad;
    if (sidx < 0 || sidx >= P.out_len) continue;
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int j = kK - 1; j >= 0; --j) {                                             // increasing frame order, like librosa.istft
      const int f = q - j;
      if (f >= 0 && f < nv) {
        const float2 v = bufs[f * PB + off[j]];
        a0 += v.y * w[j].x;
        a1 += v.x * w[j].y;
      }
    }
    const int T_lo = t0 + q - kK + 1, T_hi = t0 + q;
    const bool whole = (T_lo >= t0 || T_lo <= 0 && t0 == 0) && (T_hi < t0 + FT || t0 + FT >= P.n_frames);
    float* o = orow + sidx;
    const float v0 = a0 * P.inv_wss[sidx];
    if (whole) *o = v0; else atomicAdd(o, v0);
    if (sidx + 1 < P.out_len) {
      const float v1 = a1 * P.inv_wss[sidx + 1];
      if (whole) o[1] = v1; else atomicAdd(o + 1, v1);
    }
  }
}

// kOla = false: the windowed frames go to the workspace (istft_generic_ola_kernel adds them up afterwards).
// kOla = true:  the tile's frames are overlap-added straight from their shared-memory buffers (thread = output sample, a gather
//               over the <= N / hop frames that cover it, in increasing frame order like librosa.istft), scaled by 1 / wss and
//               written once; the <= N - hop samples at either end that the neighbouring tile also reaches are combined with
//               one atomic add each into the zeroed output -- two contributions per sample (the launcher guarantees
//               N <= (FT + 1) hop), and 0 + a + b == 0 + b + a, so the result does not depend on the order the tiles finish in.
template <int kN, bool kOla>
__global__ void __launch_bounds__(kP2Threads, 2) istft_pow2_kernel(const Pow2InvParams G) {
  extern __shared__ __align__(16) float smem[];
  constexpr Pow2Geom g = pow2_geom(kN);
  const InvParams& P = G.P;
  const P2Smem sm = p2_smem(g, smem);
  const float2* tw = sm.tw;
  const float* win = sm.win;
  float2* bufs = sm.bufs;
  p2_fill_tables(g, P.window, sm, 1.0f / (float)g.M);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const P2Slots s = p2_slots(g, lane);
  float2* sf = bufs + warp * g.fps * g.PB;
  for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const int b = tile / P.tiles_per_clip, t0 = (tile - b * P.tiles_per_clip) * g.FT;
    const bool db = P.db_flags ? (P.db_flags[b] != 0) : false;
    {
      const int f = threadIdx.x & (g.FT - 1), k0 = threadIdx.x >> g.logFT, kstep = kP2Threads >> g.logFT;
      const int t = t0 + f;
      float2* z = bufs + f * g.PB;
      const long long base = (long long)b * G.F * P.T + t;
      if (t >= P.n_frames) {
        for (int k = k0; k <= g.M; k += kstep) z[p2_pad(k)] = make_float2(0.0f, 0.0f);
      } else if (P.spec && P.gl_mag) {      // Griffin-Lim: the phase update of librosa.griffinlim inside the load (InvLoadGL's arithmetic)
        const float2* __restrict__ reb = P.spec + base;
        const float2* __restrict__ prv = P.gl_prev + base;
        const float* __restrict__ mg = P.gl_mag + base;
        const float alpha = P.gl_alpha;
#pragma unroll 4
        for (int k = k0; k <= g.M; k += kstep) {
          const long long o = (long long)k * P.T;
          const float2 r = __ldg(reb + o), t2 = __ldg(prv + o);
          const float m = __ldg(mg + o);
          const float ax = r.x - alpha * t2.x, ay = r.y - alpha * t2.y;
          const float sc = fast_div(1.0f, fast_sqrt(ax * ax + ay * ay) + kFltMin);
          z[p2_pad(k)] = make_float2((ax * sc) * m, (ay * sc) * m);
        }
      } else if (P.spec) {             // complex input: nothing but loads, eight in flight per thread
        const float2* __restrict__ src = P.spec + base + (long long)k0 * P.T;
        if (kstep >= 16) {             // kstep is a multiple of 16: the padded index advances by a constant, like the source row
          const long long rs = (long long)kstep * P.T;
          float2* zz = z + p2_pad(k0);
#pragma unroll kP2LoadUnroll
          for (int i = 0; i < g.M / kstep; ++i, src += rs, zz += kstep + kstep / 16) *zz = __ldg(src);
          if (k0 == 0) z[p2_pad(g.M)] = __ldg(P.spec + base + (long long)g.M * P.T);
        } else {
#pragma unroll 8
          for (int k = k0; k <= g.M; k += kstep) z[p2_pad(k)] = __ldg(P.spec + base + (long long)k * P.T);
        }
      } else {
#pragma unroll 4
        for (int k = k0; k <= g.M; k += kstep) {
          float xr, xi;
          inv_load_runtime(P, base + (long long)k * P.T, db, xr, xi);
          z[p2_pad(k)] = make_float2(xr, xi);
        }
      }
    }
    __syncthreads();
    p2_split_inv(g, sf, tw, lane);
    {
      const P2SmemLoad load{sf + s.fA * g.PB, sf + s.fB * g.PB};
      p2_pass_first(g, s, sf, load);
    }
    p2_passes_rest(g, s, sf, sm, lane);
    // natural order, swapped: .y = x[2n] * M, .x = x[2n + 1] * M
    if (!kOla) {
      for (int f = 0; f < g.fps; ++f) {
        const int t = t0 + warp * g.fps + f;
        if (t >= P.n_frames) break;
        const float2* z = sf + f * g.PB;
        float2* dst = reinterpret_cast<float2*>(G.frames + ((long long)b * P.T + t) * g.N);
        for (int n = lane; n < g.M; n += 32) {
          const float2 v = z[p2_pad(n)], w2 = *reinterpret_cast<const float2*>(win + 2 * n);
          dst[n] = make_float2(v.y * w2.x, v.x * w2.y);
        }
      }
    } else {
      __syncthreads();
      const int nv = (P.n_frames - t0) < g.FT ? (P.n_frames - t0) : g.FT;          // live frames of this tile
      const int span = (nv - 1) * P.hop + g.N;
      const int p0 = t0 * P.hop;                                                     // padded-signal position of the tile
      float* orow = P.out + (long long)b * P.out_pitch;
      if (G.ola_k > 0) {
        switch (G.ola_k) {
          case 1: p2_gather_pow2hop<1>(P, bufs, win, g.PB, g.FT, G.log_hop, t0, nv, span, orow); break;
          case 2: p2_gather_pow2hop<2>(P, bufs, win, g.PB, g.FT, G.log_hop, t0, nv, span, orow); break;
          case 4: p2_gather_pow2hop<4>(P, bufs, win, g.PB, g.FT, G.log_hop, t0, nv, span, orow); break;
          default: p2_gather_pow2hop<8>(P, bufs, win, g.PB, g.FT, G.log_hop, t0, nv, span, orow); break;
        }
        __syncthreads();
        continue;
      }
      // a thread takes the sample pair (u, u + 1), u even, when the hop is even (both samples then lie in the same frames
      // and in one float2 of each), else single samples.  Frames covering tile position u: f hop <= u < f hop + N.
      const int step = (P.hop & 1) ? 1 : 2;
      for (int u = step * threadIdx.x; u < span; u += step * kP2Threads) {
        const int sidx = p0 + u - P.pad;
        if (sidx + step <= 0 || sidx >= P.out_len) continue;
        // (divisions by the hop through its reciprocal, InvParams::hop_magic: exact for the tile-local operands, pow2_ola_ok)
        const int q = magic_div(u, P.hop_magic);                                  // last frame that starts at or before u
        const int below = u >= g.N ? magic_div(u - g.N, P.hop_magic) + 1 : 0;     // first frame that still reaches u
        const int f_hi = q < nv - 1 ? q : nv - 1;
        float a0 = 0.0f, a1 = 0.0f;
        int n = u - below * P.hop;
        const float2* zf = bufs + below * g.PB;
        for (int f = below; f <= f_hi; ++f, n -= P.hop, zf += g.PB) {
          const float2 v = zf[p2_pad(n >> 1)];                                    // (.x, .y) = samples (n | 1, n & ~1)
          if (step == 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(win + n);
            a0 += v.y * w2.x;
            a1 += v.x * w2.y;
          } else {
            a0 += ((n & 1) ? v.x : v.y) * win[n];
          }
        }
        // frames of the CLIP that reach this position: [T_lo, T_hi]; all of them in this tile -> the sample is complete
        const int T_lo = u >= g.N ? t0 + below : t0 - magic_div(g.N - u - 1, P.hop_magic);      // may be negative: clip start
        const int T_hi = t0 + q;                                                   // may exceed n_frames - 1: clip end
        const bool whole = (T_lo >= t0 || T_lo <= 0 && t0 == 0) && (T_hi < t0 + g.FT || t0 + g.FT >= P.n_frames);
        float* o = orow + sidx;
        if (sidx >= 0) {
          const float val = a0 * P.inv_wss[sidx];
          if (whole) *o = val; else atomicAdd(o, val);
        }
        if (step == 2 && sidx + 1 < P.out_len) {
          const float val = a1 * P.inv_wss[sidx + 1];
          if (whole) o[1] = val; else atomicAdd(o + 1, val);
        }
      }
    }
    __syncthreads();
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------------------
bool pow2_ok(int n_fft) { return tunables().pow2 != 0 && is_pow2(n_fft) && n_fft >= 64 && n_fft <= 2048; }

static size_t pow2_smem(const Pow2Geom& g) {
  return ((size_t)2 * g.M + g.N + kP2TabFloats) * sizeof(float) + (size_t)g.FT * g.PB * sizeof(float2);
}

// one instantiation per n_fft
#define AIP_P2_SIZES(X) X(64) X(128) X(256) X(512) X(1024) X(2048)

cudaError_t launch_fwd_pow2(FwdParams P, int n_fft, const DevInfo& di, cudaStream_t st) {
  Pow2FwdParams G;
  const Pow2Geom g = pow2_geom(n_fft);
  G.F = g.M + 1;
  P.tiles_per_clip = (P.T_out + g.FT - 1) / g.FT;
  if ((long long)P.B * P.tiles_per_clip > 0x7fffffffLL) return cudaErrorInvalidValue;
  P.n_tiles = (int)((long long)P.B * P.tiles_per_clip);
  G.vec = ((P.hop & 1) == 0) && ((P.pad & 1) == 0) && ((P.wave_pitch & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.wave) & 7) == 0);
  // stage the tile's span when the bulk copy is legal (16-byte granules on both sides) and two CTAs per SM still fit
  size_t smem = pow2_smem(g);
  const long long span = (long long)(g.FT - 1) * P.hop + n_fft;
  const bool vec16 = ((P.hop & 3) == 0) && ((P.pad & 3) == 0) && ((P.wave_pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.wave) & 15) == 0);
  G.span = 0;
  if (tunables().pow2_span && vec16 && 2 * (smem + (size_t)span * sizeof(float) + 1024) <= (size_t)di.max_smem_sm) {
    G.span = (int)span;
    smem += (size_t)span * sizeof(float);
  }
  G.P = P;
  void (*kern)(Pow2FwdParams) = nullptr;
  switch (n_fft) {
#define X(n) case n: kern = stft_pow2_fwd_kernel<n>; break;
    AIP_P2_SIZES(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long grid = (long long)di.sms * 2;
  if (grid > P.n_tiles) grid = P.n_tiles;
  kern<<<(unsigned)grid, kP2Threads, smem, st>>>(G);
  return cudaGetLastError();
}

// the inverse writes the waveform itself (no workspace) when at most two tiles reach any output sample
bool pow2_ola_ok(int n_fft, int hop) {
  if (!pow2_ok(n_fft) || hop <= 0) return false;
  const Pow2Geom g = pow2_geom(n_fft);
  // ... and the tile-local positions (< FT hop + N) stay in the range where the reciprocal division is exact (< 2^32 / hop)
  return (long long)n_fft <= (long long)(g.FT + 1) * hop && ((long long)g.FT * hop + n_fft) * hop < (1LL << 32);
}

// frames == nullptr: overlap-add fused (pow2_ola_ok; P.out is zeroed here first); else the frames [B, T, N] (windowed, scaled)
// of the first P.n_frames frames of every clip go to `frames`
cudaError_t launch_inv_pow2(InvParams P, int n_fft, float* frames, const DevInfo& di, cudaStream_t st) {
  Pow2InvParams G;
  const Pow2Geom g = pow2_geom(n_fft);
  G.F = g.M + 1;
  G.frames = frames;
  P.tiles_per_clip = (P.n_frames + g.FT - 1) / g.FT;
  if ((long long)P.B * P.tiles_per_clip > 0x7fffffffLL) return cudaErrorInvalidValue;
  P.n_tiles = (int)((long long)P.B * P.tiles_per_clip);
  P.hop_magic = (unsigned)((0x100000000ULL + (unsigned)P.hop - 1) / (unsigned)P.hop);
  G.ola_k = 0; G.log_hop = 0;
  if (is_pow2(P.hop) && P.hop >= 2 && P.hop <= 512 && n_fft / P.hop <= 8 && tunables().pow2_ola_fast) {      // K = 1, 2, 4 or 8
    G.ola_k = n_fft / P.hop;
    G.log_hop = ilog2(P.hop);
  }
  G.P = P;
  void (*kern)(Pow2InvParams) = nullptr;
  switch (n_fft) {
#define X(n) case n: kern = frames ? istft_pow2_kernel<n, false> : istft_pow2_kernel<n, true>; break;
    AIP_P2_SIZES(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
  const size_t smem = pow2_smem(g);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (!frames) {
    e = cudaMemset2DAsync(P.out, (size_t)P.out_pitch * sizeof(float), 0, (size_t)P.out_len * sizeof(float), (size_t)P.B, st);
    if (e != cudaSuccess) return e;
  }
  long long grid = (long long)di.sms * 2;
  if (grid > P.n_tiles) grid = P.n_tiles;
  kern<<<(unsigned)grid, kP2Threads, smem, st>>>(G);
  return cudaGetLastError();
}

}  // namespace aip
