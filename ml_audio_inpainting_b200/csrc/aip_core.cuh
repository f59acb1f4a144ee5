// Per-thread building blocks of the n_fft = 512 spectrogram kernels.
//
// This header is compiled twice:
//   * by nvcc into the sm_100a kernels of aip_kernels.cu (AIP_HD = __device__), and
//   * by g++ into csrc/aip_emul.cpp, a lane-by-lane host replay of the very same stage
//     functions that the CPU test-suite checks against the oracle (no GPU in CI).
//
// Algorithm (forward; the inverse runs the same graph backwards)
// --------------------------------------------------------------
// librosa.stft (reference call site utils.py:225-232) computes, per frame, rfft(w * x) with
// n_fft = 512.  A frame of 512 real samples is packed into 256 complex values
// z[m] = x[2m] + j x[2m+1]; the 256-point complex FFT is factored 16 x 16:
//     m = n1 + 16 n2,  k = k2 + 16 k1
//     stage 1 (16 lanes per frame, lane = n1): Y[n1][k2] = sum_n2 z[n1+16 n2] W16^(n2 k2),  times W256^(n1 k2)
//     stage 2 (lane = frame, warp = k2 pair):  Zc[k2+16 k1] = sum_n1 Y'[n1][k2] W16^(n1 k1)
// and the real spectrum follows from the split pass
//     X[k]     =      (E - jT),   X[256-k] = conj(E + jT),
//     E = Zc[k] + conj Zc[256-k],  T = W512^k (Zc[k] - conj Zc[256-k])        (the 1/2 is folded into the window).
// A stage-2 thread owns jobs (p, 16-p) (or (0, 8)), i.e. both members of every (k, 256-k) pair, so
// the split pass needs no further exchange, and since its lane index is the FRAME index every
// global store of bin k is 32 consecutive floats of row k of the [F, T] output (T contiguous, the
// layout librosa returns) -- no transposition through shared memory.
//
// Shared memory per tile of 32 frames: the staged waveform (31*hop + 512 floats, every sample
// fetched from HBM once although frames overlap 2.67x / 4x) and one exchange buffer of
// 16*16*33 float2 (index (k2*16 + n1)*33 + frame: conflict-free for both the lane = n1 writers and
// the lane = frame readers).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define AIP_HD __device__ __forceinline__
#define AIP_HDX __host__ __device__ __forceinline__
#define AIP_HM __device__ __forceinline__
#define AIP_TABLE __constant__
#else
#define AIP_HD static inline
#define AIP_HDX static inline
#define AIP_HM inline
#define AIP_TABLE static const
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace aip {

#include "aip_twiddles.inc"

constexpr int kNfft = 512;
constexpr int kBins = 257;
constexpr int kFR = 32;                    // frames per tile
constexpr int kXP = 33;                    // exchange-buffer pitch in float2 (odd: conflict-free)
constexpr int kExch = 16 * 16 * kXP;       // float2 elements in the exchange buffer

// ---- magnitude kinds of the forward epilogue (mirrors the callers' numpy epilogues) -------------
enum MagKind : int {
  MAG_NONE = 0,
  MAG_ABS = 1,          // |S|                                   models/CNNBLSTM/dataset.py:103
  MAG_LOG10_EPS = 2,    // log10(|S| + eps)                      models/CNNBLSTM/dataset.py:106
  MAG_LOG1P_POW = 3,    // log1p(|S| ** power)                   models/GAN/dataset.py:121-122
  MAG_POW = 4           // |S| ** power                          models/GAN/dataset.py:121 (normalize off)
};
// ---- magnitude domains accepted by the inverse prologue ------------------------------------------
enum MagDomain : int {
  DOM_LINEAR = 0,
  DOM_POW10 = 1,        // 10 ** x                               models/model_eval.py:163
  DOM_DB = 2,           // 10 ** (x / 20)  librosa.db_to_amplitude, utils.py:313-314
  DOM_EXPM1 = 3         // expm1(x)  (inverse of the GAN log1p)
};

// W16^m for the in-register radix-4 x radix-4 transform
constexpr float kC1 = 0.92387953251128674f;   // cos(pi/8)
constexpr float kS1 = 0.38268343236508977f;   // sin(pi/8)
constexpr float kR2 = 0.70710678118654752f;   // sqrt(1/2)

// slot that holds output bin k after fft16() (4x4 index transpose)
AIP_HD constexpr int perm16(int k) { return ((k & 3) << 2) | (k >> 2); }

AIP_HD void radix4(float& ar, float& ai, float& br, float& bi,
                   float& cr, float& ci, float& dr, float& di) {
  const float s0r = ar + cr, s0i = ai + ci, s1r = ar - cr, s1i = ai - ci;
  const float s2r = br + dr, s2i = bi + di, s3r = br - dr, s3i = bi - di;
  ar = s0r + s2r; ai = s0i + s2i;          // y0
  cr = s0r - s2r; ci = s0i - s2i;          // y2
  br = s1r + s3i; bi = s1i - s3r;          // y1 = s1 - j s3
  dr = s1r - s3i; di = s1i + s3r;          // y3 = s1 + j s3
}

// x *= (wr + j wi)
AIP_HD void cmul(float& xr, float& xi, float wr, float wi) {
  const float tr = xr * wr - xi * wi;
  xi = xr * wi + xi * wr;
  xr = tr;
}

// Forward 16-point complex DFT, in place; input natural order, output bin k in slot perm16(k).
// The inverse (unnormalised, e^{+j}) is fft16(im, re).
AIP_HD void fft16(float (&r)[16], float (&i)[16]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
    radix4(r[a], i[a], r[a + 4], i[a + 4], r[a + 8], i[a + 8], r[a + 12], i[a + 12]);
  // slot a + 4q now holds y_q of column a; twiddle by W16^(a q)
  cmul(r[5], i[5], kC1, -kS1);                                   // W16^1
  cmul(r[9], i[9], kR2, -kR2);                                   // W16^2
  cmul(r[13], i[13], kS1, -kC1);                                 // W16^3
  cmul(r[6], i[6], kR2, -kR2);                                   // W16^2
  { const float t = r[10]; r[10] = i[10]; i[10] = -t; }          // W16^4 = -j
  cmul(r[14], i[14], -kR2, -kR2);                                // W16^6
  cmul(r[7], i[7], kS1, -kC1);                                   // W16^3
  cmul(r[11], i[11], -kR2, -kR2);                                // W16^6
  cmul(r[15], i[15], -kC1, kS1);                                 // W16^9
#pragma unroll
  for (int q = 0; q < 4; ++q)
    radix4(r[4 * q], i[4 * q], r[4 * q + 1], i[4 * q + 1], r[4 * q + 2], i[4 * q + 2], r[4 * q + 3], i[4 * q + 3]);
}

// Per-lane constants of the stage that runs with lane = n1 (forward stage 1, inverse stage B):
// the 32 window taps this lane touches and the 16 inter-stage twiddles W256^(n1 k2).
struct LaneConst {
  float we[16], wo[16];
  float twr[16], twi[16];
};

// window: n_fft centre-padded taps (float); scale: 0.5 forward (split pass), 1/512 inverse (irfft norm)
AIP_HD void lane_const_init(LaneConst& c, const float* window, int n1, float scale) {
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    c.we[n2] = window[2 * (n1 + 16 * n2)] * scale;
    c.wo[n2] = window[2 * (n1 + 16 * n2) + 1] * scale;
    const float2 t = kTw256[(n1 * n2) & 255];   // n2 plays the role of k2 here
    c.twr[n2] = t.x;
    c.twi[n2] = t.y;
  }
}

// ---------------------------------------------------------------------------------------------------
// Forward, stage 1.  One call = one (frame, n1) job: 16 strided float2 loads of the staged waveform,
// window, 16-point DFT over n2, inter-stage twiddle, 16 float2 stores into the exchange buffer.
// ---------------------------------------------------------------------------------------------------
AIP_HD void fwd_stage1(const float* tile, float2* exch, int hop, int f, int n1, const LaneConst& c) {
  float r[16], i[16];
  const float* src = tile + f * hop + 2 * n1;
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const float2 z = *reinterpret_cast<const float2*>(src + 32 * n2);
    r[n2] = z.x * c.we[n2];
    i[n2] = z.y * c.wo[n2];
  }
  fft16(r, i);
  float2* dst = exch + n1 * kXP + f;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    float xr = r[perm16(k2)], xi = i[perm16(k2)];
    if (k2 > 0) cmul(xr, xi, c.twr[k2], c.twi[k2]);
    dst[k2 * 16 * kXP] = make_float2(xr, xi);
  }
}

// Split-pass twiddles of one pair-job, register resident (a stage-2 / stage-A warp keeps its job p for
// the whole kernel).  p > 0: w[k1] = W512^(p + 16 k1).  p = 0: w[k1] = W512^(8 + 16 k1) for k1 < 8 (job 8)
// and w[8 + k1] = W512^(16 k1), k1 = 1..7 (job 0).
struct PairTw {
  float wr[16], wi[16];
};

AIP_HD void pair_tw_init(PairTw& w, int p) {
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    const int k = (p != 0) ? p + 16 * k1 : (k1 < 8 ? 8 + 16 * k1 : 16 * (k1 - 8));
    const float2 t = kTw512[k];
    w.wr[k1] = t.x;
    w.wi[k1] = t.y;
  }
}

// split pass for one (k, 256-k) pair: Zk = Zc[k], Zn = Zc[256-k], (wr, wi) = W512^k.
// Output rows are addressed through the emitter's two cursors: lo(j) = bin k_lo + 16 j, hi(j) = bin k_hi - 16 j.
template <class Emit>
AIP_HD void fwd_pair(float zkr, float zki, float znr, float zni, float wr, float wi, int j, Emit& emit) {
  const float er = zkr + znr, ei = zki - zni;
  const float orr = zkr - znr, oi = zki + zni;
  const float tr = orr * wr - oi * wi;
  const float ti = orr * wi + oi * wr;
  emit.lo(j, er + ti, ei - tr);
  emit.hi(j, er - ti, -(ei + tr));
}

// ---------------------------------------------------------------------------------------------------
// Forward, stage 2.  One (frame f, pair-job p) job, p = 0..7, in two steps so that the exchange buffer
// can be handed back to the stage-1 warps as soon as it has been read:
//   fwd_stage2_load     32 float2 of jobs (p, 16-p) or (0, 8) -> registers
//   fwd_stage2_compute  two 16-point DFTs over n1, the split pass, 32 (p>0) / 33 (p=0) bins to the emitter
// ---------------------------------------------------------------------------------------------------
AIP_HD void fwd_stage2_load(const float2* exch, int f, int p, float (&ar)[16], float (&ai)[16],
                            float (&br)[16], float (&bi)[16]) {
  const int ja = p, jb = (p == 0) ? 8 : 16 - p;
  const float2* sa = exch + ja * 16 * kXP + f;
  const float2* sb = exch + jb * 16 * kXP + f;
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const float2 a = sa[n1 * kXP];
    const float2 b = sb[n1 * kXP];
    ar[n1] = a.x; ai[n1] = a.y; br[n1] = b.x; bi[n1] = b.y;
  }
}

template <class Emit>
AIP_HD void fwd_stage2_compute(float (&ar)[16], float (&ai)[16], float (&br)[16], float (&bi)[16],
                               const PairTw& w, int p, Emit& emit) {
  fft16(ar, ai);
  fft16(br, bi);
  if (p != 0) {
    emit.rows(p, 256 - p);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1)
      fwd_pair(ar[perm16(k1)], ai[perm16(k1)], br[perm16(15 - k1)], bi[perm16(15 - k1)], w.wr[k1], w.wi[k1], k1, emit);
  } else {
    emit.rows(0, 256);
    const float z0r = ar[perm16(0)], z0i = ai[perm16(0)];
    emit.lo(0, 2.0f * (z0r + z0i), 0.0f);
    emit.hi(0, 2.0f * (z0r - z0i), 0.0f);
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1)
      fwd_pair(ar[perm16(k1)], ai[perm16(k1)], ar[perm16(16 - k1)], ai[perm16(16 - k1)], w.wr[8 + k1], w.wi[8 + k1], k1, emit);
    emit.lo(8, 2.0f * ar[perm16(8)], -2.0f * ai[perm16(8)]);
    emit.rows(8, 248);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1)
      fwd_pair(br[perm16(k1)], bi[perm16(k1)], br[perm16(15 - k1)], bi[perm16(15 - k1)], w.wr[k1], w.wi[k1], k1, emit);
  }
}

// ---------------------------------------------------------------------------------------------------
// Inverse, stage A.  One call = one (frame f, pair-job p) job: fetch the bins it owns through the
// loader's cursors (lo(j) = bin k_lo + 16 j, hi(j) = bin k_hi - 16 j), undo the split pass, two inverse
// 16-point DFTs over k1, raw results to the exchange buffer (index (k2*16 + n1)*33 + f).
// scipy.fft.irfft drops imag(DC) and imag(Nyquist); so do we.
// ---------------------------------------------------------------------------------------------------
AIP_HD void inv_pair(float akr, float aki, float bkr, float bki, float wr, float wi,
                     float& zkr, float& zki, float& znr, float& zni) {
  // A = X[k], B = X[256-k], (wr, wi) = W512^k
  const float er = akr + bkr, ei = aki - bki;
  const float orr = akr - bkr, oi = aki + bki;
  const float tr = orr * wr + oi * wi;       // conj(W) * O
  const float ti = oi * wr - orr * wi;
  zkr = er - ti; zki = ei + tr;
  znr = er + ti; zni = tr - ei;
}

template <class Load, class BeforeStore>
AIP_HD void inv_stageA(float2* exch, const PairTw& w, int f, int p, bool live, Load& load, BeforeStore& before_store) {
  float ar[16], ai[16], br[16], bi[16];
  const int ja = p, jb = (p == 0) ? 8 : 16 - p;
  if (live) {
    if (p != 0) {
      load.rows(p, 256 - p);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        float xr, xi, yr, yi;
        load.lo(k1, xr, xi);
        load.hi(k1, yr, yi);
        inv_pair(xr, xi, yr, yi, w.wr[k1], w.wi[k1], ar[k1], ai[k1], br[15 - k1], bi[15 - k1]);
      }
    } else {
      float xr, xi, yr, yi;
      load.rows(0, 256);
      load.lo(0, xr, xi);
      load.hi(0, yr, yi);
      ar[0] = xr + yr; ai[0] = xr - yr;
#pragma unroll
      for (int k1 = 1; k1 < 8; ++k1) {
        load.lo(k1, xr, xi);
        load.hi(k1, yr, yi);
        inv_pair(xr, xi, yr, yi, w.wr[8 + k1], w.wi[8 + k1], ar[k1], ai[k1], ar[16 - k1], ai[16 - k1]);
      }
      load.lo(8, xr, xi);
      ar[8] = 2.0f * xr; ai[8] = -2.0f * xi;
      load.rows(8, 248);
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) {
        load.lo(k1, xr, xi);
        load.hi(k1, yr, yi);
        inv_pair(xr, xi, yr, yi, w.wr[k1], w.wi[k1], br[k1], bi[k1], br[15 - k1], bi[15 - k1]);
      }
    }
    fft16(ai, ar);      // inverse transform: swapped roles
    fft16(bi, br);
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) { ar[q] = ai[q] = br[q] = bi[q] = 0.0f; }
  }
  before_store();      // the exchange buffer must have been released by the previous tile's readers
  float2* da = exch + ja * 16 * kXP + f;
  float2* db = exch + jb * 16 * kXP + f;
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    da[n1 * kXP] = make_float2(ar[perm16(n1)], ai[perm16(n1)]);
    db[n1 * kXP] = make_float2(br[perm16(n1)], bi[perm16(n1)]);
  }
}

// ---------------------------------------------------------------------------------------------------
// Inverse, stage B.  One call = one (frame f, n1) job: conj inter-stage twiddle, inverse 16-point DFT
// over k2, synthesis window (carrying the 1/512 of irfft), windowed samples written IN PLACE into the
// exchange buffer, which thereby becomes the frame buffer: float2 slot (m*33 + f) = samples (2m, 2m+1)
// of frame f, m = n1 + 16 n2.
// ---------------------------------------------------------------------------------------------------
AIP_HD void inv_stageB(float2* exch, int f, int n1, const LaneConst& c) {
  float r[16], i[16];
  float2* p = exch + n1 * kXP + f;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float2 u = p[k2 * 16 * kXP];
    float xr = u.x, xi = u.y;
    if (k2 > 0) cmul(xr, xi, c.twr[k2], -c.twi[k2]);
    r[k2] = xr; i[k2] = xi;
  }
  fft16(i, r);
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2)
    p[n2 * 16 * kXP] = make_float2(r[perm16(n2)] * c.we[n2], i[perm16(n2)] * c.wo[n2]);
}

// Geometry of an inverse tile: 32 frames are computed, FO of them worth of output hops are produced;
// HL / HH = number of earlier / later halo frames that overlap the tile's output samples.
struct InvGeom {
  int hop, pad, HL, HH, FO;
};
AIP_HDX InvGeom inv_geom(int hop, int pad) {
  InvGeom g;
  g.hop = hop; g.pad = pad;
  g.HL = (kNfft - 1 - pad) / hop;
  g.HH = pad > 0 ? (pad - 1) / hop : -1;
  g.FO = kFR - 1 - g.HL - g.HH;
  return g;
}

// floor(n / d) for 0 <= n < 2^32 / d via a host-precomputed reciprocal m = ceil(2^32 / d)
AIP_HD int magic_div(int n, unsigned m) {
#if defined(__CUDACC__)
  return (int)__umulhi((unsigned)n, m);
#else
  return (int)(((unsigned long long)(unsigned)n * m) >> 32);
#endif
}

// Overlap-add of output sample PAIRS out of the frame buffer (float2 slot (n/2)*33 + fl holds samples
// (n, n+1) of local frame fl).  A pair sits at offset r (even) of local frame h, i.e. the frames covering it
// are fl = h - m with n = r + m*hop < 512, m = 0 .. K-1, K = ceil(512 / hop); only the m = K-1 term can be
// missing.  librosa adds frames in increasing order = decreasing m.

// interior tiles (all 32 local frames exist), K known at compile time: all K loads are independent
template <int K>
AIP_HD float2 ola_pair_fixed(const float2* fbuf, int h, int r, int hop) {
  const int step = (hop >> 1) * kXP - 1;            // slot(m) - slot(m-1)
  const int n_top = r + (K - 1) * hop;
  const float2* src = fbuf + (r >> 1) * kXP + h;     // m = 0
  float2 v[K];
#pragma unroll
  for (int m = 0; m < K - 1; ++m) v[m] = src[m * step];
  v[K - 1] = (n_top < kNfft) ? src[(K - 1) * step] : make_float2(0.0f, 0.0f);
  float sx = 0.0f, sy = 0.0f;
#pragma unroll
  for (int m = K - 1; m >= 0; --m) { sx += v[m].x; sy += v[m].y; }
  return make_float2(sx, sy);
}

// general form: any K, and clip-edge tiles where only local frames [fl_min, fl_max] exist
AIP_HD float2 ola_pair(const float2* fbuf, int h, int r, int hop, int K, int fl_min, int fl_max) {
  const int step = (hop >> 1) * kXP - 1;
  int m = K - 1;
  int n = r + m * hop;
  const float2* src = fbuf + (n >> 1) * kXP + (h - m);
  float sx = 0.0f, sy = 0.0f;
  for (; m >= 0; --m, n -= hop, src -= step) {
    const int fl = h - m;
    if (n < kNfft && fl >= fl_min && fl <= fl_max) { const float2 v = *src; sx += v.x; sy += v.y; }
  }
  return make_float2(sx, sy);
}

}  // namespace aip
