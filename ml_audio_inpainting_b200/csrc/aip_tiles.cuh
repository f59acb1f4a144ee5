// Tile-level bodies of the n_fft = 512 kernels, written as "phase functions": a kernel is
//     for (tile ...) { phase0(tid); sync; phase1(tid); sync; phase2(tid); sync; }
// so that csrc/aip_emul.cpp can replay exactly the same code on the host, thread by thread
// (for phase: for tid: body), and the CPU test-suite can check indexing and arithmetic against
// the oracle without a GPU.  See aip_core.cuh for the algorithm.
#pragma once
#include "aip_core.cuh"

namespace aip {

constexpr int kThreads = 256;              // 8 warps: stage 2 has 8 pair-jobs per frame
constexpr float kLog10of2 = 0.30102999566398120f;
constexpr float kLog2of10 = 3.32192809488736235f;
constexpr float kFltMin = 1.17549435e-38f; // np.finfo(float32).tiny

#if defined(__CUDACC__)
AIP_HD float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
AIP_HD float fast_log2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
AIP_HD float fast_exp2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
AIP_HD void fast_sincos(float a, float& s, float& c) { __sincosf(a, &s, &c); }
#else
AIP_HD float fast_sqrt(float x) { return sqrtf(x); }
AIP_HD float fast_log2(float x) { return log2f(x); }
AIP_HD float fast_exp2(float x) { return exp2f(x); }
AIP_HD void fast_sincos(float a, float& s, float& c) { s = sinf(a); c = cosf(a); }
#endif
#if defined(__CUDACC__)
AIP_HD float fast_div(float a, float b) { return __fdividef(a, b); }
#else
AIP_HD float fast_div(float a, float b) { return a / b; }
#endif

// np.angle / atan2 for the phase outputs (models/GAN/dataset.py:123, models/model_eval.py:86): one division, a degree-7
// polynomial in a^2 on [0, 1] (fitted for this file; max error 1.3e-7 rad in fp32, i.e. half an ulp of pi) and three
// sign / quadrant selects -- about a third of atan2f's instructions in an epilogue that runs once per bin.
AIP_HD float fast_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = mx > 0.0f ? fast_div(mn, mx) : 0.0f;
  const float s = a * a;
  float p = -0.004054083023220301f;
  p = p * s + 0.021861141547560692f;
  p = p * s - 0.05590960383415222f;
  p = p * s + 0.09641990065574646f;
  p = p * s - 0.13908545672893524f;
  p = p * s + 0.19946548342704773f;
  p = p * s - 0.33329859375953674f;
  p = p * s + 0.9999993443489075f;
  float r = p * a;
  if (ay > ax) r = 1.57079632679489662f - r;
  if (copysignf(1.0f, x) < 0.0f) r = 3.14159265358979324f - r;      // sign bit, so that atan2(+0, -0) = pi like IEEE / numpy
  return copysignf(r, y);
}

// log1p(m), m >= 0 (models/GAN/dataset.py:122,135): log(u) * m / (u - 1) with u = fl(1 + m) cancels the rounding of u
// expm1 for the GAN back-end's un-log (the inverse of its front-end's log1p): ex2.approx - 1 away from zero (relative error of the
// difference <= 5 x 2^-22 for |x| >= 0.25), a degree-6 Taylor polynomial near it (remainder x^6 / 5040 < 5e-8 relative).  Measured
// against expm1 in float64 over [-20, 20]: <= 1.2e-6 relative (tests/test_emul_kernels.py); libdevice's expm1f cost the mag + phase
// inverse 0.24 ms of 0.62.
AIP_HD float fast_expm1(float x) {
  if (fabsf(x) >= 0.25f) return fast_exp2(x * 1.44269504088896341f) - 1.0f;
  const float p = fmaf(x, fmaf(x, fmaf(x, fmaf(x, 1.0f / 720.0f, 1.0f / 120.0f), 1.0f / 24.0f), 1.0f / 6.0f), 0.5f);
  return fmaf(x * x, p, x);
}

AIP_HD float fast_log1p(float m) {
  const float u = 1.0f + m, d = u - 1.0f;
  const float r = fast_log2(u) * 0.693147180559945309f;
  return d == 0.0f ? m : r * fast_div(m, d);
}

// ===================================================================================================
// Forward
// ===================================================================================================
struct FwdParams {
  const float* wave;        // [B, L], row pitch wave_pitch (elements)
  long long wave_pitch;
  int B, L;
  int hop, pad;             // pad = n_fft/2 when center else 0  (librosa.stft zero padding)
  int T;                    // frames per clip = 1 + (L + 2 pad - 512) / hop
  int T_out;                // frames written (crop, models/CNNBLSTM/dataset.py:109-111), <= T
  const float* window;      // [512] centre-padded analysis window
  const int* gap_samples;   // [B,2] or null: zero samples [g0,g1) before the transform (utils.py:141-142,180-183)
  const int* zero_frames;   // [B,2] or null: zero the SPECTRUM on frames [f0,f1) (models/model_eval.py:154)
  const int* mask_frames;   // [B,2] or null: frame range of the dense mask
  int mask_in_gap_is_one;   // 1: CNNBLSTM convention (dataset.py:115-118); 0: GAN convention (GAN/dataset.py:150-152)
  int mag_kind;
  float eps, power;
  float2* spec;             // [B,257,T_out] or null
  float* mag;               // [B,257,T_out] or null
  float* phase;             // [B,257,T_out] or null
  float* mask;              // [B,257,T_out] or null
  int tiles_per_clip;
  int n_tiles;              // B * tiles_per_clip (< 2^31)
  int tile_floats;          // floats per staged-waveform buffer (tile length rounded up to 128 B)
  int n_tile_bufs;          // 2: the next tile's copy overlaps stage 1; 1: large hops
  int var_div;              // gap-variant mode (FWD_VARIANT, aip_stft_gap_variants_f32), else 0: "clip" b is variant b of
                            // wave row b / var_div (gaps_per_audio, models/CNNBLSTM/dataset.py:93) and only the tiles_per_clip
                            // tiles starting at frame var_frame_base(P, gap start of b) are transformed
  int chunk;                // tiles per draw from the dynamic schedule
  unsigned* tile_counter;   // device counter of the dynamic schedule, zero at launch
  int reflect;              // centre padding by reflection (librosa < 0.10's default pad_mode; aip_stft_desc.center == 2), else zeros
  int vec_ok;               // 16-byte bulk copies of the waveform are legal (alignment)
  int zero_groups;          // win_zero_groups(win_length): the first / last 32 * zero_groups window taps are zero
  int var_prefetch;         // gap-variant mode: request the lines of a tile's row segments from L2 ahead of its stores
  const int* var_meta;      // gap-variant mode: [B, 4] = {gap start, gap end, var_frame_base, wave row} per variant, written by
                            // variant_meta_kernel (no divisions or dependent index loads inside the tile loops)
};

AIP_HDX int fwd_tile_len(int hop) { return (kFR - 1) * hop + kNfft; }

// Where a tile's samples come from.  Tile element i <-> clip sample g0 + i.  Elements
// [v_lo, v_lo + n_bulk) are moved by ONE 1-D TMA bulk copy (cp.async.bulk, 16-byte granules) issued a
// whole phase ahead; the rest (zero padding outside [0, L), a <= 3 element ragged tail, the gap range)
// is patched by the threads once the copy has landed (fwd_fixup).
struct FwdTilePlan {
  const float* src;     // clip base
  int g0;               // clip sample of tile element 0 (may be negative: centre padding)
  int len;              // tile elements
  int v_lo, v_hi;       // tile elements that map to real samples: [v_lo, v_hi)
  int n_bulk;           // elements moved by the bulk copy (multiple of 4), starting at v_lo
  int gs, ge;           // gap range in clip samples
  int L, reflect;       // clip length; padding outside [0, L) by reflection instead of zeros
};

// A CTA owns a contiguous run of tiles; (clip b, tile-in-clip tt) advance without divisions.
struct TileCursor {
  int b, tt;
};
AIP_HD TileCursor tile_cursor(int tix, int tiles_per_clip) {
  TileCursor c;
  c.b = tix / tiles_per_clip;
  c.tt = tix - c.b * tiles_per_clip;
  return c;
}
AIP_HD void tile_advance(TileCursor& c, int tiles_per_clip) {
  if (++c.tt == tiles_per_clip) { c.tt = 0; ++c.b; }
}

AIP_HD FwdTilePlan fwd_tile_plan(const FwdParams& P, const TileCursor& c) {
  FwdTilePlan q;
  q.len = fwd_tile_len(P.hop);
  q.g0 = c.tt * kFR * P.hop - P.pad;
  q.src = P.wave + (long long)c.b * P.wave_pitch;
  q.v_lo = q.g0 < 0 ? -q.g0 : 0;
  const int hi = P.L - q.g0;
  q.v_hi = hi < q.len ? (hi > 0 ? hi : 0) : q.len;
  if (q.v_lo > q.v_hi) q.v_lo = q.v_hi;
  q.n_bulk = P.vec_ok ? ((q.v_hi - q.v_lo) & ~3) : 0;
  q.gs = q.ge = 0;
  q.L = P.L; q.reflect = P.reflect;
  if (P.gap_samples) { q.gs = P.gap_samples[2 * c.b]; q.ge = P.gap_samples[2 * c.b + 1]; }
  return q;
}

// The same plan with the clip's gap range supplied by the caller: the persistent kernel fetches it from global memory
// once per CLIP, not once per tile (the load's latency sat in front of every tile's fix-up test).
AIP_HD FwdTilePlan fwd_tile_plan_gap(const FwdParams& P, const TileCursor& c, int gs, int ge) {
  FwdTilePlan q;
  q.len = fwd_tile_len(P.hop);
  q.g0 = c.tt * kFR * P.hop - P.pad;
  q.src = P.wave + (long long)c.b * P.wave_pitch;
  q.v_lo = q.g0 < 0 ? -q.g0 : 0;
  const int hi = P.L - q.g0;
  q.v_hi = hi < q.len ? (hi > 0 ? hi : 0) : q.len;
  if (q.v_lo > q.v_hi) q.v_lo = q.v_hi;
  q.n_bulk = P.vec_ok ? ((q.v_hi - q.v_lo) & ~3) : 0;
  q.gs = gs; q.ge = ge;
  q.L = P.L; q.reflect = P.reflect;
  return q;
}

// ---- gap variants (models/CNNBLSTM/dataset.py:93-111 computes gaps_per_audio full STFTs per file) ----
// A gap [gs, ge) changes only the frames whose 512-sample span [t hop - pad, t hop - pad + 512) meets it:
//   t in [ floor((gs + pad - 512) / hop) + 1 , ceil((ge + pad) / hop) ).
// A variant recomputes tiles_per_clip tiles of kFR frames starting at the frame below; every other frame is a copy of
// the clean spectrogram.  The base is pulled back so that the run ends at T_out (all tiles full) where it would overshoot.
AIP_HDX int var_first_frame(int gs, int pad, int hop) {
  const int a = gs + pad - kNfft;                 // floor(a / hop) + 1, a may be negative
  return a < 0 ? 0 : a / hop + 1;
}
AIP_HDX int var_tiles(int gap_len_max, int hop, int T_out) {
  // frames met by a gap of <= gap_len_max samples: at most ceil((gap_len_max + 512) / hop) + 1
  const int n_aff = (gap_len_max + kNfft + hop - 1) / hop + 1;
  const int need = (n_aff + kFR - 1) / kFR, all = (T_out + kFR - 1) / kFR;
  return need < all ? need : all;
}
AIP_HD int var_frame_base(const FwdParams& P, int gs) {
  int fb = var_first_frame(gs, P.pad, P.hop);
  const int last = P.T_out - P.tiles_per_clip * kFR;
  if (fb > last) fb = last;
  return fb < 0 ? 0 : fb;
}
AIP_HD FwdTilePlan fwd_tile_plan_var(const FwdParams& P, const TileCursor& c, int gs, int ge, int fb, int row) {
  FwdTilePlan q;
  q.len = fwd_tile_len(P.hop);
  q.g0 = (fb + c.tt * kFR) * P.hop - P.pad;
  q.src = P.wave + (long long)row * P.wave_pitch;
  q.v_lo = q.g0 < 0 ? -q.g0 : 0;
  const int hi = P.L - q.g0;
  q.v_hi = hi < q.len ? (hi > 0 ? hi : 0) : q.len;
  if (q.v_lo > q.v_hi) q.v_lo = q.v_hi;
  // the bulk copy needs a 16-byte aligned source: g0 + v_lo is a multiple of 4 when hop, pad are (vec_ok) and fb is any int
  q.n_bulk = P.vec_ok ? ((q.v_hi - q.v_lo) & ~3) : 0;
  q.gs = gs; q.ge = ge;
  q.L = P.L; q.reflect = P.reflect;
  return q;
}

AIP_HD bool fwd_needs_fixup(const FwdTilePlan& q) {
  return q.v_lo > 0 || q.v_lo + q.n_bulk < q.len || (q.ge > q.g0 && q.gs < q.g0 + q.len && q.ge > q.gs);
}

// everything of the tile that the bulk copy does not deliver
// Every element is written by at most ONE thread here: the scalar copy applies the gap itself (a separate zeroing
// pass by other threads would race with it -- seen on hardware with unaligned waveforms), the zeroing pass below only
// touches what the bulk copy delivered.
// element i of the tile lies outside the clip: zero (librosa >= 0.10, pad_mode="constant") or the GAPPED clip mirrored about its
// first / last sample (np.pad(..., mode="reflect"), librosa < 0.10); positions past the mirror's reach belong to frames that do
// not exist and read as zero
AIP_HD float fwd_pad_value(const FwdTilePlan& q, int i) {
  if (!q.reflect) return 0.0f;
  int g = q.g0 + i;
  g = g < 0 ? -g : 2 * (q.L - 1) - g;
  if (g < 0 || g >= q.L || (g >= q.gs && g < q.ge)) return 0.0f;
  return q.src[g];
}

AIP_HD void fwd_fixup(const FwdTilePlan& q, int tid, float* tile) {
  int a = q.gs - q.g0, e = q.ge - q.g0;        // gap range (utils.py:141-142, :180-183) in tile elements
  if (a < 0) a = 0;
  if (e > q.len) e = q.len;
  for (int i = tid; i < q.v_lo; i += kThreads) tile[i] = fwd_pad_value(q, i);
  const int bulk_end = q.v_lo + q.n_bulk;
  for (int i = bulk_end + tid; i < q.v_hi; i += kThreads) tile[i] = (i >= a && i < e) ? 0.0f : q.src[q.g0 + i];
  for (int i = q.v_hi + tid; i < q.len; i += kThreads) tile[i] = fwd_pad_value(q, i);
  if (a < q.v_lo) a = q.v_lo;
  if (e > bulk_end) e = bulk_end;
  for (int i = a + tid; i < e; i += kThreads) tile[i] = 0.0f;
}

// Gap zeroing WITHOUT a CTA-wide barrier (gap-variant mode, where every tile holds a gap): a stage-1 warp reads the frames
// {2w, 2w + 1, 2w + 16, 2w + 17} of the tile (fwd_phase1), so it zeroes the gap inside the sample spans of exactly those
// frames itself and needs a __syncwarp() only.  Frames of different warps overlap: such elements are zeroed by every warp
// that reads them (identical values), and a warp never reads a gap element it has not zeroed.  With one barrier per tile the
// stage-1 warps ran in lockstep with the elected thread's request path (ncu: 18 % of the samples on that barrier).
AIP_HDX bool fwd_needs_edge_fixup(const FwdTilePlan& q) { return q.v_lo > 0 || q.v_lo + q.n_bulk < q.len; }
AIP_HD void fwd_gap_zero_own(const FwdTilePlan& q, int hop, int tid, float* tile) {
  int a = q.gs - q.g0, e = q.ge - q.g0;
  if (a < 0) a = 0;
  if (e > q.len) e = q.len;
  const int warp = tid >> 5, lane = tid & 31;
  for (int half = 0; half < 2; ++half) {
    const int f = 2 * warp + 16 * half;
    int lo = f * hop, hi = (f + 1) * hop + kNfft;
    if (lo < a) lo = a;
    if (hi > e) hi = e;
    for (int i = lo + lane; i < hi; i += 32) tile[i] = 0.0f;
  }
}

// stage 1 of the FFT for one tile: 256 threads, lane = n1 (16 lanes per frame); a thread owns column n1 of
// frames fa and fa + 16 and runs them as the two lanes of the packed FP32x2 codelet
template <int ZP, class Before>
AIP_HD void fwd_phase1(const FwdParams& P, int tid, const float* tile, float2* exch, const float* win_s,
                       const LaneConst& lc, Before& before_store) {
  const int warp = tid >> 5, lane = tid & 31;
  const int fa = 2 * warp + (lane >> 4);
  fwd_stage1<ZP>(tile, exch, win_s, P.hop, fa, fa + 16, lane & 15, lc, before_store);
}
template <int ZP>
AIP_HD void fwd_phase1(const FwdParams& P, int tid, const float* tile, float2* exch, const float* win_s,
                       const LaneConst& lc) {
  NoWait nw;
  fwd_phase1<ZP>(P, tid, tile, exch, win_s, lc, nw);
}

// zero n2 groups (32 taps) at each end of the centre-padded window that the kernels exploit: 2 or 0
AIP_HDX int win_zero_groups(int win_length) {
  return (win_length > 0 && win_length <= kNfft - 128) ? 2 : 0;
}

AIP_HD float mag_value(int mk, float xr, float xi, float eps, float power) {
  if (mk == MAG_POW && power == 2.0f) return xr * xr + xi * xi;      // the power spectrogram (librosa.feature.melspectrogram's default)
  const float m = fast_sqrt(xr * xr + xi * xi);
  if (mk == MAG_ABS) return m;
  if (mk == MAG_LOG10_EPS) return fast_log2(m + eps) * kLog10of2;
  const float mp = (power == 1.0f) ? m : powf(m, power);
  return (mk == MAG_LOG1P_POW) ? fast_log1p(mp) : mp;
}

// two bins at once on FP32x2 (power == 1; MAG_POW: power == 2); the MUFU ops stay scalar
AIP_HD float2 mag_value2(int mk, float2 xr, float2 xi, float eps) {
  const float2 pw = fma2(xi, xi, mul2(xr, xr));
  if (mk == MAG_POW) return pw;
  const float2 m = make_float2(fast_sqrt(pw.x), fast_sqrt(pw.y));
  if (mk == MAG_ABS) return m;
  if (mk == MAG_LOG10_EPS) {
    const float2 l = add2(m, make_float2(eps, eps));
    return mul2s(make_float2(fast_log2(l.x), fast_log2(l.y)), kLog10of2);
  }
  return (mk == MAG_LOG1P_POW) ? make_float2(fast_log1p(m.x), fast_log1p(m.y)) : m;
}

// Forward kernel variants: a bit mask of what the epilogue produces, a template parameter of the kernel so
// that each variant is straight-line code (no per-bin branches, no calls).
//   bits 0..2  magnitude kind (MagKind: 0 none, 1 |S|, 2 log10(|S|+eps), 3 log1p(|S|), 4 |S|**2)
//   FWD_SPEC   complex output          FWD_PHASE  angle(S)        FWD_MASK  dense frame mask
//   FWD_ZERO   spectrum-domain gap (frames [f0,f1) zeroed before the epilogue, models/model_eval.py:154)
//   FWD_FULL   run-time flags, out-of-line epilogue: every other combination (|S|**p with p != 1, ...)
enum FwdModeBits : int { FWD_SPEC = 8, FWD_PHASE = 16, FWD_MASK = 32, FWD_ZERO = 64, FWD_FULL = 128, FWD_VARIANT = 256 };
constexpr int FWD_MAG_ABS = MAG_ABS, FWD_MAG_LOG10 = MAG_LOG10_EPS;      // the two magnitude-only variants

// A stage-2 thread writes bins k_lo + 16 j and k_hi - 16 j of ONE frame (its lane): two row offsets are set
// per job (rows()), lo(j) / hi(j) give the element offset of a bin from the column pointer -- and, the lane being
// the frame index, a warp's store of one bin is 32 consecutive elements of that row.  No predicates: in the last
// tile of a clip the lanes past T_out recompute the last valid frame and store the same values.
// put2<SX, SY> takes two bins as a packed pair whose imaginary parts still carry the signs (SX, SY) (see
// fwd_pair2); magnitudes ignore them, complex / phase outputs apply them.
// kT > 0: T_out is known at compile time (the reference's fixed shapes: 417 frames for 5 s, 834 for 10 s at hop 192), so a
// bin's offset from its row cursor is an immediate of the store instead of an IMAD + IMAD.WIDE per store.
template <int kM, int kT = 0>
struct FwdEmitT {
  typedef int off_t;
  float* mag;               // column pointers: array + b*F*T_out + t
  float2* spec;
  float* phase;
  float* mask;
  int T;                    // T_out
  float eps, maskv;
  bool zero;
  int olo, ohi, s16;
  AIP_HM int tv() const { return kT > 0 ? kT : T; }
  AIP_HM void rows(int k_lo, int k_hi) { olo = k_lo * tv(); ohi = k_hi * tv(); s16 = 16 * tv(); }
  AIP_HM int lo(int j) const { return olo + j * (16 * tv()); }
  AIP_HM int hi(int j) const { return ohi - j * (16 * tv()); }
  AIP_HM int bin(int k) const { return k * tv(); }
  AIP_HM void put1(int o, float xr, float xi) const {
    if (kM & FWD_ZERO) { if (zero) { xr = 0.0f; xi = 0.0f; } }
    if (kM & FWD_SPEC) spec[o] = make_float2(xr, xi);
    if (kM & FWD_PHASE) phase[o] = fast_atan2(xi, xr);
    if (kM & FWD_MASK) mask[o] = maskv;
    if (kM & 7) mag[o] = mag_value(kM & 7, xr, xi, eps, (kM & 7) == MAG_POW ? 2.0f : 1.0f);
  }
  template <int SX, int SY>
  AIP_HM void put2(int ox, int oy, float2 xr, float2 xi) const {
    if (kM & FWD_ZERO) { if (zero) { xr = make_float2(0.0f, 0.0f); xi = make_float2(0.0f, 0.0f); } }
    if (kM & (FWD_SPEC | FWD_PHASE)) {
      const float ix = SX < 0 ? -xi.x : xi.x, iy = SY < 0 ? -xi.y : xi.y;
      if (kM & FWD_SPEC) { spec[ox] = make_float2(xr.x, ix); spec[oy] = make_float2(xr.y, iy); }
      if (kM & FWD_PHASE) { phase[ox] = fast_atan2(ix, xr.x); phase[oy] = fast_atan2(iy, xr.y); }
    }
    if (kM & FWD_MASK) { mask[ox] = maskv; mask[oy] = maskv; }
    if (kM & 7) {
#if defined(AIP_MAG_SCALAR)
      mag[ox] = mag_value(kM & 7, xr.x, xi.x, eps, (kM & 7) == MAG_POW ? 2.0f : 1.0f);
      mag[oy] = mag_value(kM & 7, xr.y, xi.y, eps, (kM & 7) == MAG_POW ? 2.0f : 1.0f);
#else
      const float2 m = mag_value2(kM & 7, xr, xi, eps);
#if defined(AIP_ABLATE_STORES)     // timing experiment only: everything is computed, (almost) nothing is written
      if (m.x == 123.456f) mag[ox] = m.x;
      if (m.y == 123.456f) mag[oy] = m.y;
#else
      mag[ox] = m.x;
      mag[oy] = m.y;
#endif
#endif
    }
  }
};

// Epilogue, general path: any mix of complex / magnitude / phase / mask outputs, the spectrum-domain gap and |S| ** p.
// The emitter carries COPIES of the few FwdParams fields it needs: holding a reference to the kernel parameter block made it
// addressable (a 272-byte local-memory copy per thread, read back on every store) and the out-of-line store routine it used to
// call 65 times per job kept half the register file alive across each call -- together 8.3 ms where two specialised launches
// producing the same outputs take 0.86 ms.  Now the stores are inline behind CTA-uniform tests and only the rare transcendental
// tails (|S| ** p for p other than 1 and 2, via powf) stay out of line.
#if defined(__CUDACC__)
static __device__ __noinline__
#else
static inline
#endif
float mag_value_general(int mk, float xr, float xi, float eps, float power) { return mag_value(mk, xr, xi, eps, power); }
// every magnitude flavour that needs no powf: |S|, log10(|S| + eps), log1p(|S|), |S| ** 1, |S| ** 2 (`square`)
AIP_HD float mag_value_nopow(int mk, float xr, float xi, float eps, bool square) {
  const float pw = xr * xr + xi * xi;
  if (mk == MAG_POW && square) return pw;
  const float m = fast_sqrt(pw);
  if (mk == MAG_LOG10_EPS) return fast_log2(m + eps) * kLog10of2;
  if (mk == MAG_LOG1P_POW) return fast_log1p(m);
  return m;
}

// ... and two bins at once on FP32x2
AIP_HD float2 mag_value2_nopow(int mk, float2 xr, float2 xi, float eps, bool square) {
  const float2 pw = fma2(xi, xi, mul2(xr, xr));
  if (mk == MAG_POW && square) return pw;
  const float2 m = make_float2(fast_sqrt(pw.x), fast_sqrt(pw.y));
  if (mk == MAG_LOG10_EPS) {
    const float2 l = add2(m, make_float2(eps, eps));
    return mul2s(make_float2(fast_log2(l.x), fast_log2(l.y)), kLog10of2);
  }
  return (mk == MAG_LOG1P_POW) ? make_float2(fast_log1p(m.x), fast_log1p(m.y)) : m;
}

// FwdEmitT with its template switches as CTA-uniform run-time tests: same column pointers, int offsets and packed pairs
struct FwdEmitFull {
  typedef int off_t;
  float* mag;               // column pointers (array + b*F*T_out + t), null = not produced
  float2* spec;
  float* phase;
  float* mask;
  int T, mag_kind;          // T = T_out
  float eps, power, maskv;
  bool active, zero, slow_mag, square;
  int olo, ohi, s16;
  AIP_HM void rows(int k_lo, int k_hi) { olo = k_lo * T; ohi = k_hi * T; s16 = 16 * T; }
  AIP_HM int lo(int j) const { return olo + j * s16; }
  AIP_HM int hi(int j) const { return ohi - j * s16; }
  AIP_HM int bin(int k) const { return k * T; }
  AIP_HM void put1(int o, float xr, float xi) const {
    if (!active) return;
    if (zero) { xr = 0.0f; xi = 0.0f; }
    if (spec) spec[o] = make_float2(xr, xi);
    if (phase) phase[o] = fast_atan2(xi, xr);
    if (mask) mask[o] = maskv;
    if (mag) mag[o] = slow_mag ? mag_value_general(mag_kind, xr, xi, eps, power) : mag_value_nopow(mag_kind, xr, xi, eps, square);
  }
  template <int SX, int SY>
  AIP_HM void put2(int ox, int oy, float2 xr, float2 xi) const {
    if (!active) return;
    if (zero) { xr = make_float2(0.0f, 0.0f); xi = make_float2(0.0f, 0.0f); }
    if (spec || phase) {
      const float ix = SX < 0 ? -xi.x : xi.x, iy = SY < 0 ? -xi.y : xi.y;
      if (spec) { spec[ox] = make_float2(xr.x, ix); spec[oy] = make_float2(xr.y, iy); }
      if (phase) { phase[ox] = fast_atan2(ix, xr.x); phase[oy] = fast_atan2(iy, xr.y); }
    }
    if (mask) { mask[ox] = maskv; mask[oy] = maskv; }
    if (mag) {
      if (slow_mag) {
        mag[ox] = mag_value_general(mag_kind, xr.x, xi.x, eps, power);
        mag[oy] = mag_value_general(mag_kind, xr.y, xi.y, eps, power);
      } else {
        const float2 m = mag_value2_nopow(mag_kind, xr, xi, eps, square);
        mag[ox] = m.x;
        mag[oy] = m.y;
      }
    }
  }
};

// (b, t): the frame the emitter's column pointers address; the row offsets k * T_out must fit an int (the launchers check)
AIP_HD FwdEmitFull fwd_make_emit_full(const FwdParams& P, int b, int t, int n_bins, bool active) {
  FwdEmitFull emit;
  const long long col = (long long)b * n_bins * P.T_out + t;
  emit.mag = (P.mag_kind != MAG_NONE && P.mag) ? P.mag + col : nullptr;
  emit.spec = P.spec ? P.spec + col : nullptr;
  emit.phase = P.phase ? P.phase + col : nullptr;
  emit.mask = P.mask ? P.mask + col : nullptr;
  emit.T = P.T_out; emit.mag_kind = P.mag_kind; emit.eps = P.eps; emit.power = P.power;
  emit.active = active; emit.zero = false; emit.maskv = 0.0f;
  emit.olo = emit.ohi = emit.s16 = 0;
  // powers the straight-line magnitude code does not cover (it knows 1, and 2 for MAG_POW)
  emit.slow_mag = (P.mag_kind == MAG_POW && P.power != 2.0f && P.power != 1.0f) || (P.mag_kind == MAG_LOG1P_POW && P.power != 1.0f);
  emit.square = P.power == 2.0f;
  if (P.zero_frames) emit.zero = (t >= P.zero_frames[2 * b] && t < P.zero_frames[2 * b + 1]);
  if (P.mask) {
    bool in = false;
    if (P.mask_frames) in = (t >= P.mask_frames[2 * b] && t < P.mask_frames[2 * b + 1]);
    emit.maskv = (in == (P.mask_in_gap_is_one != 0)) ? 1.0f : 0.0f;
  }
  return emit;
}

struct NoRelease { AIP_HM void operator()() const {} };

// the variants that are instantiated; anything else runs FWD_FULL
AIP_HDX bool fwd_mode_is_fast(int m) {
  return m == FWD_MAG_ABS || m == FWD_MAG_LOG10 || m == FWD_SPEC || m == (MAG_LOG10_EPS | FWD_MASK) ||
         m == MAG_LOG1P_POW || m == (MAG_LOG1P_POW | FWD_PHASE | FWD_MASK) || m == (FWD_SPEC | FWD_PHASE | FWD_MASK) ||
         m == (MAG_LOG10_EPS | FWD_ZERO) || m == (MAG_ABS | FWD_PHASE) || m == (MAG_LOG1P_POW | FWD_PHASE) ||
         m == (FWD_SPEC | FWD_PHASE) || m == MAG_POW ||
         // "the spectrogram and its magnitude" in one launch (what a caller asks for first; __graft_entry__.smoke does)
         m == (FWD_SPEC | MAG_ABS) || m == (FWD_SPEC | MAG_LOG10_EPS) || m == (FWD_SPEC | MAG_LOG1P_POW) ||
         m == (MAG_LOG10_EPS | FWD_PHASE);
}

AIP_HDX int fwd_mode_of(const FwdParams& P) {
  int m = P.mag_kind;
  // |S| ** p: the power spectrogram p == 2 (the mel front-end, utils.py:268-277) is a variant of its own, any other p runs FWD_FULL
  if ((P.mag_kind == MAG_POW && P.power != 2.0f) || (P.mag_kind == MAG_LOG1P_POW && P.power != 1.0f)) return FWD_FULL;
  if (P.spec) m |= FWD_SPEC;
  if (P.phase) m |= FWD_PHASE;
  if (P.mask) m |= FWD_MASK;
  if (P.zero_frames) m |= FWD_ZERO;
  return fwd_mode_is_fast(m) ? m : (int)FWD_FULL;
}

// stage 2 + split pass + epilogue for one tile: 256 threads, lane = frame, warp = pair-job.
// `release` runs once the exchange buffer has been read into registers.
template <int kMode, class Release, int kT = 0>
AIP_HD void fwd_phase2(const FwdParams& P, int tid, const TileCursor& c, const float2* exch, const PairTw& w,
                       Release& release, int var_fb = 0) {
  const int p = tid >> 5, lane = tid & 31;
  int t0 = c.tt * kFR;
  if (kMode & FWD_VARIANT) t0 += var_fb;      // var_frame_base(P, gap start of variant c.b), fetched by the caller
  const int T_out = kT > 0 ? kT : P.T_out;    // kT: compile-time T_out (the launcher guarantees P.T_out == kT)
  const int n_valid = (T_out - t0) < kFR ? (T_out - t0) : kFR;
  const int fr = lane < n_valid ? lane : n_valid - 1;     // lanes past the end replay the last valid frame
  float2 zr[16], zi[16];
  fwd_stage2_load(exch, fr, p, zr, zi);
#ifndef AIP_FWD_LATE_RELEASE
#define AIP_FWD_LATE_RELEASE 5
#endif
  // When is the exchange buffer handed back to the stage-1 warps?  As soon as it has been read, where stage 1 is the slower
  // role (complex output alone).  In every variant that computes magnitudes or phases stage 2 is the slower role and stage 1 has
  // time to spare; handed the buffer early, stage 1 of the tile after next starts at once and competes for the FP32 pipe with this
  // tile's two DFTs, which are FP32-pipe bound -- every change that made stage 1 faster made these variants SLOWER.  There the
  // buffer is handed back only after AIP_FWD_LATE_RELEASE - 1 of the 8 split-pass / epilogue rounds (MUFU, stores: the FP32
  // pipe has room).  Headline launch, A/B on one box: release after the loads 1.465 ms, after the DFTs 1.448, after 2 / 3 / 4
  // rounds 1.430 / 1.433 / 1.430, after 6 / 8 rounds 1.52 / 1.56 (stage 1 starts too late).  Since stage 1 transforms its tile before
  // it waits for the buffer (aip_fwd.cu) the release sits after 4 rounds: 1.410 -> 1.388 ms.  The GAN front-end (log1p + phase +
  // mask, win 512 / hop 128, 1024 x 5 s): 1.513 -> 1.433 ms; the eval front-end (complex + phase, then log10 + spectrum gap): 0.890 -> 0.839.
  constexpr bool kLate = AIP_FWD_LATE_RELEASE && (kMode & (7 | FWD_PHASE)) != 0 && kMode != FWD_FULL;
  if (!kLate) release();
  NoHook no_hook;
  const long long col = (long long)c.b * kBins * T_out + t0 + fr;
  if (kMode == FWD_FULL) {
    FwdEmitFull emit = fwd_make_emit_full(P, c.b, t0 + fr, kBins, lane < n_valid);
    fwd_stage2_compute(zr, zi, w, p, emit, no_hook);
  } else {
    const int t = t0 + fr;
    FwdEmitT<kMode, kT> emit{(kMode & 7) ? P.mag + col : nullptr, (kMode & FWD_SPEC) ? P.spec + col : nullptr,
                             (kMode & FWD_PHASE) ? P.phase + col : nullptr, (kMode & FWD_MASK) ? P.mask + col : nullptr,
                             T_out, P.eps, 0.0f, false, 0, 0, 0};
    if (kMode & FWD_MASK) {
      bool in = false;
      if (P.mask_frames) in = (t >= P.mask_frames[2 * c.b] && t < P.mask_frames[2 * c.b + 1]);
      emit.maskv = (in == (P.mask_in_gap_is_one != 0)) ? 1.0f : 0.0f;
    }
    if (kMode & FWD_ZERO) emit.zero = (t >= P.zero_frames[2 * c.b] && t < P.zero_frames[2 * c.b + 1]);
    if (kLate) fwd_stage2_compute<FwdEmitT<kMode, kT>, Release, AIP_FWD_LATE_RELEASE - 2>(zr, zi, w, p, emit, release);
    else fwd_stage2_compute(zr, zi, w, p, emit, no_hook);
  }
}

// ===================================================================================================
// Inverse
// ===================================================================================================
struct InvParams {
  const float2* spec;       // [B,257,T] complex, or null
  const float* mag;         // [B,257,T] magnitude (with mag_domain), used when spec is null
  const float* phase;       // [B,257,T] or null (zero phase)
  int mag_domain;
  const int* db_flags;      // [B] or null: per clip, non-zero => treat mag as dB (utils.py:313-314 heuristic)
  const float* blend_in;    // [B,257,T] or null: mag := mag * blend_mask + blend_in * (1 - blend_mask) first
  const float* blend_mask;  // [B,257,T]        (StackedBLSTMCNN.reconstruct_spectrogram, models/CNNBLSTM/model.py:108)
  int B, T;                 // T = frames per clip in the input (row pitch)
  int n_frames;             // frames used (librosa.istft: min(T, ceil(padded_length / hop)) when length given)
  int hop, pad;
  const float* window;      // [512] synthesis window
  const float* inv_wss;     // [out_len]: 1/window_sumsquare where > tiny else 1 (librosa.istft normalisation)
  int out_len;
  float* out;               // [B, out_len], row pitch out_pitch
  float* peaks;             // [B] or null: running max |sample| per clip (zeroed by the caller; fused peak normalisation)
  long long out_pitch;
  int vec_ok;               // float2 loads of inv_wss / stores of the output are legal
  int tiles_per_clip;
  int n_tiles;
  int tiles_per_cta;
  int n_bufs;               // exchange buffers in the ring (1..kInvBufs)
  int l2_prefetch;          // stage A requests the next tile's row segments from L2 while it works on this one
  InvGeom g;
  unsigned hop_magic;       // ceil(2^32 / hop)
  unsigned col_magic;       // ceil(2^32 / (hop / 2))
  int ola_terms;            // ceil(512 / hop): frames overlapping one sample
  int ola_dq, ola_dr;       // (2 * 256) / hop and (2 * 256) % hop: per-iteration advance of a thread's pair
  int wss_ref;              // inv_wss[wss_ref + r] = periodic value for frame offset r; < 0: no interior
  int ola_fast;             // compile-time specialised overlap-add of interior tiles: 0 none, 1 = hop 192 with the 2 x 64 zero
                            // taps of a win <= 384 window (2 terms), 2 = hop 128, full window (4 terms); centre padding only
  // Griffin-Lim (INV_GL): `spec` is the spectrum the forward kernel rebuilt in the previous iteration and the phase update
  // of librosa.griffinlim runs inside the load: a = spec - gl_alpha * gl_prev; a *= gl_mag / (|a| + tiny)
  const float2* gl_prev;    // [B,257,T] rebuilt spectrum of the iteration before (first iteration: = spec, with gl_alpha 0)
  const float* gl_mag;      // [B,257,T] target magnitudes; non-null selects INV_GL
  float gl_alpha;           // momentum / (1 + momentum)

};

// Prologue loaders.  A stage-A thread reads bins k_lo + 16 j and k_hi - 16 j of ONE frame (its lane), so a
// warp's load of one bin is 32 consecutive elements of that row of the [F, T] input.
// Magnitude (+ phase) input with the prologue fused, specialised at compile time.
//   kDom 0: linear, or dB where the clip's auto flag says so (utils.py:313-314)   1: 2**(x*scale) (10**x or dB)
//   kDom 2: expm1(x)
template <int kDom, bool kPhase, bool kBlend>
struct InvLoadMag {
  const float* mag;         // column pointers: array + b*F*T + t
  const float* phase;
  const float* bin;         // blend input / mask (kBlend)
  const float* bmask;
  int T;
  float scale;              // log2(10) for 10**x, log2(10)/20 for dB
  bool db;                  // kDom == 0: this clip is in dB
  int olo, ohi, s16;
  AIP_HM void rows(int k_lo, int k_hi) { olo = k_lo * T; ohi = k_hi * T; s16 = 16 * T; }
  AIP_HM void done() const {}
  AIP_HM void lo(int j, float& xr, float& xi) const { get(olo + j * s16, xr, xi); }
  AIP_HM void hi(int j, float& xr, float& xi) const { get(ohi - j * s16, xr, xi); }
  AIP_HM void get(int o, float& xr, float& xi) const {
    float m = mag[o];
    if (kBlend) { const float g = bmask[o]; m = m * g + bin[o] * (1.0f - g); }
    if (kDom == 0) m = db ? fast_exp2(m * (kLog2of10 * 0.05f)) : m;
    else if (kDom == 1) m = fast_exp2(m * scale);
    else m = fast_expm1(m);
    if (kPhase) {
      float s, c;
      fast_sincos(phase[o], s, c);
      xr = m * c; xi = m * s;
    } else { xr = m; xi = 0.0f; }
  }
};

#if defined(__CUDACC__)
AIP_HD float2 ldg_stream(const float2* p) { return __ldcg(p); }      // L2 only: the spectrum is read once
#else
AIP_HD float2 ldg_stream(const float2* p) { return *p; }
#endif

struct InvLoadSpec {        // complex input straight from HBM
  const float2* col;        // spec + b*F*T + t
  int T;
  const char* plo;
  const char* phi;
  unsigned s16b;            // bytes between rows k and k + 16 (T < 2^22)
  AIP_HM void done() const {}
  AIP_HM void rows(int k_lo, int k_hi) {
    plo = reinterpret_cast<const char*>(col + k_lo * T);
    phi = reinterpret_cast<const char*>(col + k_hi * T);
    s16b = 128u * (unsigned)T;
  }
  // row address = base +- j * s16b as ONE 32 x 32 + 64 bit multiply-add (IMAD.WIDE) instead of a shift / add-with-carry chain
  AIP_HM void lo(int j, float& xr, float& xi) const {
    const float2 v = ldg_stream(reinterpret_cast<const float2*>(plo + (unsigned long long)s16b * (unsigned)j));
    xr = v.x; xi = v.y;
  }
  AIP_HM void hi(int j, float& xr, float& xi) const {
    const float2 v = ldg_stream(reinterpret_cast<const float2*>(phi + (long long)(int)s16b * (long long)(-j)));
    xr = v.x; xi = v.y;
  }
};

// Griffin-Lim: the phase update fused into the load (same arithmetic as gl_update_pp_kernel, which it replaces: the
// projected spectrum `angles` is never written to or read from HBM)
struct InvLoadGL {
  const float2* reb;        // ARRAY bases (uniform: they stay in the constant bank); the thread keeps one column index
  const float2* prev;
  const float* mag;
  long long col;            // b*F*T + t
  int T;
  float alpha;              // 0 in the first iteration (prev = reb then: branch-free)
  int olo, ohi, s16;
  AIP_HM void rows(int k_lo, int k_hi) { olo = k_lo * T; ohi = k_hi * T; s16 = 16 * T; }
  AIP_HM void done() const {}
  AIP_HM void lo(int j, float& xr, float& xi) const { get(olo + j * s16, xr, xi); }
  AIP_HM void hi(int j, float& xr, float& xi) const { get(ohi - j * s16, xr, xi); }
  AIP_HM void get(int o, float& xr, float& xi) const {
    const long long i = col + o;
    const float2 r = ldg_stream(reb + i);
    const float2 t = ldg_stream(prev + i);
    const float m = mag[i];
    const float ax = r.x - alpha * t.x, ay = r.y - alpha * t.y;
    // librosa's order: angles /= |angles| + tiny, THEN angles *= S.  (a * (S / (|a| + tiny)) is NaN for a = 0 as soon as
    // S / tiny overflows, i.e. S > 4: an exactly silent bin of the rebuilt spectrum under a non-zero target.)
    const float sc = fast_div(1.0f, fast_sqrt(ax * ax + ay * ay) + kFltMin);      // MUFU.SQRT + MUFU.RCP: ~2 ulp each
    xr = (ax * sc) * m; xi = (ay * sc) * m;
  }
};

// run-time flavoured element fetch (generic n_fft kernels only)
AIP_HD void inv_load_runtime(const InvParams& P, long long idx, bool db, float& xr, float& xi) {
  if (P.spec) { const float2 v = P.spec[idx]; xr = v.x; xi = v.y; return; }
  float m = P.mag[idx];
  if (P.blend_in) { const float g = P.blend_mask[idx]; m = m * g + P.blend_in[idx] * (1.0f - g); }
  const int dom = db ? (int)DOM_DB : P.mag_domain;
  if (dom == DOM_POW10) m = fast_exp2(m * kLog2of10);
  else if (dom == DOM_DB) m = fast_exp2(m * (kLog2of10 * 0.05f));
  else if (dom == DOM_EXPM1) m = fast_expm1(m);
  if (P.phase) { float sn, cs; fast_sincos(P.phase[idx], sn, cs); xr = m * cs; xi = m * sn; }
  else { xr = m; xi = 0.0f; }
}

// Input modes of the inverse kernel (template parameter): INV_SPEC = complex input; otherwise
// 1 + 2*kDom + kPhase for magnitude (+ phase) input; INV_BLEND = the model hand-off: blended log10 magnitude
// (out * mask + in * (1 - mask)) -> 10** -> times exp(j phase).  (A cp.async-staged variant of the complex input, one tile
// ahead through shared memory, measured 1.47x SLOWER than plain loads: 8-byte LDGSTS throttles the LSU and adds
// an LDS per element; see profiles/README.md.)
constexpr int INV_SPEC = 0;
constexpr int INV_BLEND = 7;
constexpr int INV_GL = 8;       // complex input + the Griffin-Lim phase update (InvLoadGL)
constexpr int INV_BLEND_LIN = 9;     // the hand-off with the blended magnitude used as it is (models/GAN/train.py:473-482) ...
constexpr int INV_BLEND_EXPM1 = 10;  // ... or through expm1 (the inverse of the GAN front-end's log1p)
AIP_HDX constexpr int inv_mag_mode(int dom, bool phase) { return 1 + 2 * dom + (phase ? 1 : 0); }

AIP_HDX int inv_mode_of(const InvParams& P) {
  if (P.spec) return P.gl_mag ? INV_GL : INV_SPEC;
  if (P.blend_in) return P.mag_domain == DOM_LINEAR ? INV_BLEND_LIN : (P.mag_domain == DOM_EXPM1 ? INV_BLEND_EXPM1 : INV_BLEND);
  const int dom = (P.mag_domain == DOM_LINEAR) ? 0 : (P.mag_domain == DOM_EXPM1 ? 2 : 1);
  return inv_mag_mode(dom, P.phase != nullptr);
}

constexpr int kInvBufs = 3;     // most exchange buffers the ring supports
constexpr int kInvBufsDefault = 2;   // measured 0.753 / 0.735 / 0.991 ms for 1 / 2 / 3 buffers (L2-only loads)
constexpr int kMaxWtab = 1024;  // largest hop with a shared-memory 1/wss period table

// stage A for one tile: 256 threads, lane = frame, warp = pair-job
template <int kMode, class BeforeStore>
AIP_HD void inv_phase0(const InvParams& P, int tid, const TileCursor& c, float2* exch, const PairTw& w,
                       BeforeStore& before_store) {
  const int warp = tid >> 5, lane = tid & 31;
  const int t = c.tt * P.g.FO - P.g.HL + lane;
  const bool live = (t >= 0 && t < P.n_frames);
  const long long col = (long long)c.b * kBins * P.T + t;
  if (kMode == INV_SPEC) {
    InvLoadSpec load{P.spec + col, P.T, nullptr, nullptr, 0u};
    inv_stageA(exch, w, lane, warp, live, load, before_store);
  } else if (kMode == INV_GL) {
    InvLoadGL load{P.spec, P.gl_prev, P.gl_mag, col, P.T, P.gl_alpha, 0, 0, 0};
    inv_stageA(exch, w, lane, warp, live, load, before_store);
  } else if (kMode == INV_BLEND || kMode == INV_BLEND_LIN || kMode == INV_BLEND_EXPM1) {
    constexpr int kDom = kMode == INV_BLEND ? 1 : (kMode == INV_BLEND_LIN ? 0 : 2);
    InvLoadMag<kDom, true, true> load{P.mag + col, P.phase + col, P.blend_in + col, P.blend_mask + col, P.T,
                                      P.mag_domain == DOM_DB ? kLog2of10 * 0.05f : kLog2of10, false, 0, 0, 0};
    inv_stageA(exch, w, lane, warp, live, load, before_store);
  } else {
    constexpr int kDom = (kMode - 1) >> 1;
    constexpr bool kPhase = ((kMode - 1) & 1) != 0;
    InvLoadMag<kDom, kPhase, false> load{P.mag + col, kPhase ? P.phase + col : nullptr, nullptr, nullptr, P.T,
                                  P.mag_domain == DOM_DB ? kLog2of10 * 0.05f : kLog2of10,
                                  P.db_flags ? (P.db_flags[c.b] != 0) : false, 0, 0, 0};
    inv_stageA(exch, w, lane, warp, live, load, before_store);
  }
}

// which specialised overlap-add (if any) the launch may use: decided once on the host
AIP_HDX int inv_ola_fast_kind(int hop, int pad, int win_length) {
  if (pad != kNfft / 2) return 0;
  if (hop == 192 && win_zero_groups(win_length) == 2) return 1;
  if (hop == 128) return 2;
  return 0;
}

// Interior tile: all 32 local frames exist, the whole tile lies inside the output, vector stores are legal and the
// periodic 1/window-sum-square table is present.  Only such tiles run the specialised phases.
AIP_HD bool inv_tile_interior(const InvParams& P, const TileCursor& c) {
  const int f_first = c.tt * P.g.FO - P.g.HL;
  const bool edge = f_first < 0 || P.n_frames - 1 - f_first < kFR - 1;
  const bool has_wtab = P.wss_ref >= 0 && P.hop <= 1024;
  return !edge && has_wtab && P.vec_ok && (c.tt + 1) * P.g.FO * P.hop <= P.out_len;
}

// stage B for one tile: 256 threads, lane = n1.  kFast (template parameter of the kernel) = InvParams::ola_fast.
template <int kFast>
AIP_HD void inv_phase1(const InvParams& P, int tid, const TileCursor& c, float2* exch, const float* win_s, const LaneConst& lc) {
  const int warp = tid >> 5, lane = tid & 31;
  const int fa = 2 * warp + (lane >> 4);
  if (kFast == 1) {
    inv_stageB<2>(exch, win_s, fa, fa + 16, lane & 15, lc);
    if (!inv_tile_interior(P, c)) inv_stageB_zero_pruned(exch, fa, fa + 16, lane & 15);
  } else {
    inv_stageB<0>(exch, win_s, fa, fa + 16, lane & 15, lc);
  }
}

// Overlap-add of an interior tile with everything but the thread's column known at compile time (centre padding).
// The output pair at offset u = 2 cc + 256 of local frame F0 = HL + h is the sum over e = EMIN..EMAX of frame F0 + e at
// offset u - e HOP; for (HOP 192, window 384) that is e in {0, 1} for every column (the third covering frame only
// contributes zero taps), for (HOP 128, window 512) e in {-1, 0, 1, 2}.  A thread owns pair columns pc and pc + C/2
// (conflict-free LDS.64 over consecutive pc) and walks the hops h = g, g + G, ... fully unrolled: per pair
// (EMAX - EMIN + 1) LDS.64, (EMAX - EMIN) FADD2, one FMUL2 and one 8-byte store, no integer work.
// Returns max |sample| over what this thread stored (peak tracking of the fused normalisation, utils.py:84).
template <int HOP, int EMIN, int EMAX>
AIP_HD float inv_ola_fast(const InvParams& P, int tid, int s0, const float2* fbuf, const float* wtab, float* dst) {
  constexpr int C = HOP / 2, TC = C / 2;               // pair columns per hop, thread columns
  constexpr int G = kThreads / TC;                     // hop groups walking in parallel
  constexpr int HL = (kNfft - 1 - kNfft / 2) / HOP, HH = (kNfft / 2 - 1) / HOP, FO = kFR - 1 - HL - HH;
  constexpr int NIT = (FO + G - 1) / G;
  constexpr int STEP = C * kXP - 1;                    // slot(frame F, offset n) - slot(frame F + 1, offset n - HOP)
  const int g = tid / TC, pc = tid - g * TC;
  if (g >= G) return 0.0f;
  const int u1 = 2 * pc + kNfft / 2, u2 = u1 + 2 * TC;
  const float2 nw1 = *reinterpret_cast<const float2*>(wtab + (u1 % HOP));
  const float2 nw2 = *reinterpret_cast<const float2*>(wtab + (u2 % HOP));
  const float2* src1 = fbuf + (u1 >> 1) * kXP + HL + g;
  const float2* src2 = src1 + TC * kXP;
  float* out = dst + s0 + g * HOP + 2 * pc;
  float pk = 0.0f;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    if (it * G + G <= FO || g + it * G < FO) {
      float2 a1 = src1[it * G - EMIN * STEP], a2 = src2[it * G - EMIN * STEP];     // increasing frame order, like librosa
#pragma unroll
      for (int e = EMIN + 1; e <= EMAX; ++e) {
        a1 = add2(a1, src1[it * G - e * STEP]);
        a2 = add2(a2, src2[it * G - e * STEP]);
      }
      a1 = mul2(a1, nw1);
      a2 = mul2(a2, nw2);
      *reinterpret_cast<float2*>(out + it * G * HOP) = a1;
      *reinterpret_cast<float2*>(out + it * G * HOP + 2 * TC) = a2;
      if (P.peaks) pk = fmaxf(fmaxf(pk, fmaxf(fabsf(a1.x), fabsf(a1.y))), fmaxf(fabsf(a2.x), fabsf(a2.y)));
    }
  }
  return pk;
}

// Interior tile (all 32 local frames exist, whole tile inside the output, vector stores legal, periodic 1/wss
// table present), K terms per sample.  Column-wise: a thread owns the output pairs at ONE offset 2*cc inside
// the hop and walks the hops h = g, g + G, ...  Everything that depends on the offset only -- the frame-buffer
// base slot, the 1/wss pair, whether the m = K-1 term exists -- is loop invariant; per pair that leaves K
// LDS.64, 2K FADD, 2 FMUL, one 8-byte store and two pointer increments.
template <int K>
AIP_HD float inv_ola_interior(const InvParams& P, int tid, int s0, const float2* fbuf, const float* wtab, float* dst) {
  const int hop = P.hop, C = hop >> 1;                // pair columns per hop
  const int G = C >= kThreads ? 1 : kThreads / C;     // hop groups walking in parallel
  const int g = C >= kThreads ? 0 : magic_div(tid, P.col_magic);
  if (g >= G) return 0.0f;
  float pk = 0.0f;
  const int step = C * kXP - 1;                       // slot(m) - slot(m-1)
  for (int cc = tid - g * C; cc < C; cc += kThreads) {
    const int u = 2 * cc + P.pad;                     // offset of the pair from the start of local frame HL + h
    const int d = magic_div(u, P.hop_magic);
    const int r = u - d * hop;                        // offset inside frame h' = h + HL + d
    const bool top = r + (K - 1) * hop < kNfft;
    const float2 nw = *reinterpret_cast<const float2*>(wtab + r);
    const float2* src = fbuf + (r >> 1) * kXP + (P.g.HL + d + g);
    float* out = dst + s0 + g * hop + 2 * cc;
    for (int h = g; h < P.g.FO; h += G, src += G, out += G * hop) {
      float2 v[K];
#pragma unroll
      for (int m = 0; m < K - 1; ++m) v[m] = src[m * step];
      v[K - 1] = top ? src[(K - 1) * step] : make_float2(0.0f, 0.0f);
      float sx = 0.0f, sy = 0.0f;
#pragma unroll
      for (int m = K - 1; m >= 0; --m) { sx += v[m].x; sy += v[m].y; }    // increasing frame order, like librosa
      sx *= nw.x; sy *= nw.y;
      *reinterpret_cast<float2*>(out) = make_float2(sx, sy);
      pk = fmaxf(pk, fmaxf(fabsf(sx), fabsf(sy)));
    }
  }
  return pk;
}

// per-clip peak of the fused normalisation: non-negative floats order like their bit patterns
AIP_HD void inv_peak_commit(float* peak, float pk) {
#if defined(__CUDACC__)
  const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(pk));
  if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(reinterpret_cast<unsigned*>(peak), m);
#else
  if (pk > *peak) *peak = pk;
#endif
}

// overlap-add + window-sum-square normalisation + store for one tile (256 threads): thread q, q + 256, ...
// owns output pairs (s0 + 2q, s0 + 2q + 1); its (frame, offset) coordinates advance incrementally.
// wtab: one period of 1/window-sum-square indexed by the offset r inside a frame -- valid for every sample of
// a tile whose 32 local frames all exist (window_sumsquare is exactly hop-periodic there, same float32
// accumulation order); tiles at the clip edges (or wtab == null) read the global table instead.
template <int kFast>
AIP_HD void inv_phase2(const InvParams& P, int tid, const TileCursor& c, const float2* fbuf, const float* wtab) {
  const int hop = P.hop;
  const int f_first = c.tt * P.g.FO - P.g.HL;
  const int s0 = c.tt * P.g.FO * hop;
  int fl_min = -f_first;
  if (fl_min < 0) fl_min = 0;
  int fl_max = P.n_frames - 1 - f_first;
  if (fl_max > kFR - 1) fl_max = kFR - 1;
  const bool edge = fl_min > 0 || fl_max < kFR - 1;
  float* dst = P.out + (long long)c.b * P.out_pitch;
  float pk = 0.0f;
  bool done = false;
  if (!edge && wtab && P.vec_ok && s0 + P.g.FO * hop <= P.out_len) {
    done = true;
    if (kFast == 1) pk = inv_ola_fast<192, 0, 1>(P, tid, s0, fbuf, wtab, dst);
    else if (kFast == 2) pk = inv_ola_fast<128, -1, 2>(P, tid, s0, fbuf, wtab, dst);
    else if (P.ola_terms == 3) pk = inv_ola_interior<3>(P, tid, s0, fbuf, wtab, dst);
    else if (P.ola_terms == 4) pk = inv_ola_interior<4>(P, tid, s0, fbuf, wtab, dst);
    else if (P.ola_terms == 2) pk = inv_ola_interior<2>(P, tid, s0, fbuf, wtab, dst);
    else if (P.ola_terms == 1) pk = inv_ola_interior<1>(P, tid, s0, fbuf, wtab, dst);
    else done = false;
  }
  if (done) {
    if (P.peaks) inv_peak_commit(P.peaks + c.b, pk);
    return;
  }
  const int n_pairs = (P.g.FO * hop) >> 1;
  // position of pair q relative to local frame 0: u = 2q + pad + HL*hop = h*hop + r
  const int u0 = 2 * tid + P.pad + P.g.HL * hop;
  int h = magic_div(u0, P.hop_magic);
  int r = u0 - h * hop;
  for (int q = tid; q < n_pairs; q += kThreads) {
    const int s = s0 + 2 * q;
    if (s >= P.out_len) break;
    const float2 v = ola_pair(fbuf, h, r, hop, P.ola_terms, fl_min, fl_max);
    const float y0 = v.x * P.inv_wss[s];
    dst[s] = y0;
    pk = fmaxf(pk, fabsf(y0));
    if (s + 1 < P.out_len) {
      const float y1 = v.y * P.inv_wss[s + 1];
      dst[s + 1] = y1;
      pk = fmaxf(pk, fabsf(y1));
    }
    h += P.ola_dq;
    r += P.ola_dr;
    if (r >= hop) { r -= hop; ++h; }
  }
  if (P.peaks) inv_peak_commit(P.peaks + c.b, pk);
}

}  // namespace aip
