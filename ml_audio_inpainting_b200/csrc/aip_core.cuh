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
// Every 16-point DFT is the register-resident radix-4 x radix-4 codelet fft16x2(), which runs TWO transforms
// per thread in the two lanes of Blackwell's packed FP32x2 instructions (stage 1: two frames; stage 2: the
// two jobs of a pair).  The real spectrum follows from the split pass
//     X[k]     =      (E - jT),   X[256-k] = conj(E + jT),
//     E = Zc[k] + conj Zc[256-k],  T = W512^k (Zc[k] - conj Zc[256-k])        (the 1/2 is folded into the window).
// A stage-2 thread owns jobs (p, 16-p) (or (0, 8)), i.e. both members of every (k, 256-k) pair, so
// the split pass needs no further exchange, and since its lane index is the FRAME index every
// global store of bin k is 32 consecutive floats of row k of the [F, T] output (T contiguous, the
// layout librosa returns) -- no transposition through shared memory.
//
// Shared memory per tile of 32 frames: the staged waveform (31*hop + 512 floats, every sample
// fetched from HBM once although frames overlap 2.67x / 4x) and one exchange buffer of
// 16*16*33 float2 (pitch 33 per (job-pair, re/im, n1) row: conflict-free for both the lane = n1 writers
// and the lane = frame readers; see exch_slot()).
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

// slot that holds output bin k after fft16x2() (4x4 index transpose of the radix-4 x radix-4 codelet)
AIP_HD constexpr int perm16(int k) { return ((k & 3) << 2) | (k >> 2); }

// x *= (wr + j wi)
AIP_HD void cmul(float& xr, float& xi, float wr, float wi) {
  const float tr = xr * wr - xi * wi;
  xi = xr * wi + xi * wr;
  xr = tr;
}

// ---- packed FP32x2 (Blackwell FADD2 / FMUL2 / FFMA2: two fp32 lanes per issue slot) --------------------
// The FP32 pipe does not get wider, but these kernels are issue-slot bound, and every 16-point DFT here has
// an identical twin (the two pair-jobs of a stage-2 thread, the two frames of a stage-1 thread): .x carries
// one, .y the other.
#if defined(__CUDACC__) && !defined(AIP_SCALAR_FFT)
AIP_HD float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
AIP_HD float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }   // exact a - b
AIP_HD float2 mul2s(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
AIP_HD float2 fma2s(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
AIP_HD float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
AIP_HD float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
AIP_HD float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AIP_HD float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
AIP_HD float2 mul2s(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
AIP_HD float2 fma2s(float2 a, float s, float2 c) { return make_float2(a.x * s + c.x, a.y * s + c.y); }
AIP_HD float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
AIP_HD float2 fma2(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
// half swap / full negate of a packed pair: ptxas folds both into the F32x2 operand modifiers of the consumer
// (R.F32x2.LO_HI, -R.F32x2.HI_LO), so they cost no instruction
AIP_HD float2 swap2(float2 a) { return make_float2(a.y, a.x); }
AIP_HD float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

AIP_HD void radix4x2(float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  const float2 s0r = add2(ar, cr), s0i = add2(ai, ci), s1r = sub2(ar, cr), s1i = sub2(ai, ci);
  const float2 s2r = add2(br, dr), s2i = add2(bi, di), s3r = sub2(br, dr), s3i = sub2(bi, di);
  ar = add2(s0r, s2r); ai = add2(s0i, s2i);
  cr = sub2(s0r, s2r); ci = sub2(s0i, s2i);
  br = add2(s1r, s3i); bi = sub2(s1i, s3r);
  dr = sub2(s1r, s3i); di = add2(s1i, s3r);
}

// x *= (wr + j wi), both lanes
AIP_HD void cmulx2(float2& xr, float2& xi, float wr, float wi) {
  const float2 tr = fma2s(xi, -wi, mul2s(xr, wr));
  xi = fma2s(xi, wr, mul2s(xr, wi));
  xr = tr;
}

// radix-4 butterflies with one input known to be zero (pruned window taps, see fwd_stage1<ZP>): 12 instead of 16 ops
AIP_HD void radix4x2_a0(float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  const float2 s2r = add2(br, dr), s2i = add2(bi, di), s3r = sub2(br, dr), s3i = sub2(bi, di);
  ar = add2(cr, s2r); ai = add2(ci, s2i);                       // s0 = c, s1 = -c
  br = sub2(s3i, cr); bi = sub2(neg2(ci), s3r);
  dr = sub2(neg2(cr), s3i); di = sub2(s3r, ci);
  cr = sub2(cr, s2r); ci = sub2(ci, s2i);
}
AIP_HD void radix4x2_d0(float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  const float2 s0r = add2(ar, cr), s0i = add2(ai, ci), s1r = sub2(ar, cr), s1i = sub2(ai, ci);
  ar = add2(s0r, br); ai = add2(s0i, bi);                       // s2 = s3 = b
  cr = sub2(s0r, br); ci = sub2(s0i, bi);
  dr = sub2(s1r, bi); di = add2(s1i, br);
  const float2 tr = add2(s1r, bi);
  bi = sub2(s1i, br); br = tr;
}

// second half of the 16-point codelet: W16 twiddles + the radix-4 pass over the slot groups.
// The twiddles are not applied as complex products in front of the butterflies (9 x 4 + 64 = 97 packed ops after the -j
// shortcut) but folded into them (80): every W16^m is a common real factor times (1 -+ j tan(pi/8)) or (1 -+ j), so an
// input is rotated by one FMA per component (plain add / sub for the odd multiples of W16^2), and the common factor rides
// on the FMA that forms the butterfly's output.  All factors are compile-time constants, i.e. immediates of FFMA2, which
// issues at the FADD2 rate (tools/microbench/f32x2_imm.cu).  Measured A/B on one box: forward 1.482 -> 1.467 ms, inverse
// 0.515 -> 0.502 ms; accuracy against a float64 DFT unchanged (1.1e-7 relative rms).
//   group q holds (a, b, c, d) = slots 4q .. 4q + 3, twiddles (1, W^q, W^2q, W^3q), W = W16 = e^{-j pi/8}
//   q = 1:  b W   = C1 [(br + t bi) + j (bi - t br)]     c W^2 =  R2 [(cr + ci) + j (ci - cr)]     d W^3 =  C1 [(di + t dr) + j (t di - dr)]
//   q = 2:  b W^2 = R2 [(br + bi) + j (bi - br)]         c W^4 = -j c                              d W^6 = -R2 [(dr - di) + j (dr + di)]
//   q = 3:  b W^3 = C1 [(bi + t br) + j (t bi - br)]     c W^6 = -R2 [(cr - ci) + j (cr + ci)]     d W^9 = -C1 [(dr + t di) + j (di - t dr)]
// kSkip: 0 = all four outputs, 1 = output a not needed (slot left untouched), 2 = output d not needed.
constexpr float kT1 = 0.41421356237309504880f;   // tan(pi/8)

// outputs of a radix-4 butterfly from s0, s1 (unscaled) and s2', s3' (to be scaled by f):
//   a = s0 + f s2',  c = s0 - f s2',  b = s1 - j f s3',  d = s1 + j f s3'
template <int kSkip>
AIP_HD void radix4x2_out(float f, float2 s0r, float2 s0i, float2 s1r, float2 s1i, float2 s2r, float2 s2i, float2 s3r, float2 s3i,
                         float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  if (kSkip != 1) { ar = fma2s(s2r, f, s0r); ai = fma2s(s2i, f, s0i); }
  cr = fma2s(s2r, -f, s0r); ci = fma2s(s2i, -f, s0i);
  br = fma2s(s3i, f, s1r); bi = fma2s(s3r, -f, s1i);
  if (kSkip != 2) { dr = fma2s(s3i, -f, s1r); di = fma2s(s3r, f, s1i); }
}

template <int kSkip>
AIP_HD void radix4x2_w1(float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  const float2 cpr = add2(cr, ci), cpi = sub2(ci, cr);
  const float2 s0r = fma2s(cpr, kR2, ar), s0i = fma2s(cpi, kR2, ai), s1r = fma2s(cpr, -kR2, ar), s1i = fma2s(cpi, -kR2, ai);
  const float2 bpr = fma2s(bi, kT1, br), bpi = fma2s(br, -kT1, bi);
  const float2 dpr = fma2s(dr, kT1, di), dpi = fma2s(di, kT1, neg2(dr));
  radix4x2_out<kSkip>(kC1, s0r, s0i, s1r, s1i, add2(bpr, dpr), add2(bpi, dpi), sub2(bpr, dpr), sub2(bpi, dpi),
                      ar, ai, br, bi, cr, ci, dr, di);
}
template <int kSkip>
AIP_HD void radix4x2_w2(float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  const float2 s0r = add2(ar, ci), s0i = sub2(ai, cr), s1r = sub2(ar, ci), s1i = add2(ai, cr);
  const float2 bpr = add2(br, bi), bpi = sub2(bi, br);
  const float2 dpr = sub2(dr, di), dpi = add2(dr, di);
  radix4x2_out<kSkip>(kR2, s0r, s0i, s1r, s1i, sub2(bpr, dpr), sub2(bpi, dpi), add2(bpr, dpr), add2(bpi, dpi),
                      ar, ai, br, bi, cr, ci, dr, di);
}
template <int kSkip>
AIP_HD void radix4x2_w3(float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  const float2 cpr = sub2(cr, ci), cpi = add2(cr, ci);
  const float2 s0r = fma2s(cpr, -kR2, ar), s0i = fma2s(cpi, -kR2, ai), s1r = fma2s(cpr, kR2, ar), s1i = fma2s(cpi, kR2, ai);
  const float2 bpr = fma2s(br, kT1, bi), bpi = fma2s(bi, kT1, neg2(br));
  const float2 dpr = fma2s(di, kT1, dr), dpi = fma2s(dr, -kT1, di);
  radix4x2_out<kSkip>(kC1, s0r, s0i, s1r, s1i, sub2(bpr, dpr), sub2(bpi, dpi), add2(bpr, dpr), add2(bpi, dpi),
                      ar, ai, br, bi, cr, ci, dr, di);
}

AIP_HD void fft16x2_tail(float2 (&r)[16], float2 (&i)[16]) {
  radix4x2(r[0], i[0], r[1], i[1], r[2], i[2], r[3], i[3]);
  radix4x2_w1<0>(r[4], i[4], r[5], i[5], r[6], i[6], r[7], i[7]);
  radix4x2_w2<0>(r[8], i[8], r[9], i[9], r[10], i[10], r[11], i[11]);
  radix4x2_w3<0>(r[12], i[12], r[13], i[13], r[14], i[14], r[15], i[15]);
}

// Two forward 16-point complex DFTs at once (lane .x and lane .y), in place; input natural order, output
// bin k in slot perm16(k).  The inverse (unnormalised, e^{+j}) is fft16x2(im, re).
AIP_HD void fft16x2(float2 (&r)[16], float2 (&i)[16]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
    radix4x2(r[a], i[a], r[a + 4], i[a + 4], r[a + 8], i[a + 8], r[a + 12], i[a + 12]);
  fft16x2_tail(r, i);
}

// radix-4 butterflies that skip one OUTPUT (its slot is left untouched): inverse stage B drops the samples whose
// synthesis-window tap is zero
AIP_HD void radix4x2_noA(float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  const float2 s0r = add2(ar, cr), s0i = add2(ai, ci), s1r = sub2(ar, cr), s1i = sub2(ai, ci);
  const float2 s2r = add2(br, dr), s2i = add2(bi, di), s3r = sub2(br, dr), s3i = sub2(bi, di);
  cr = sub2(s0r, s2r); ci = sub2(s0i, s2i);
  br = add2(s1r, s3i); bi = sub2(s1i, s3r);
  dr = sub2(s1r, s3i); di = add2(s1i, s3r);
}
AIP_HD void radix4x2_noD(float2& ar, float2& ai, float2& br, float2& bi, float2& cr, float2& ci, float2& dr, float2& di) {
  const float2 s0r = add2(ar, cr), s0i = add2(ai, ci), s1r = sub2(ar, cr), s1i = sub2(ai, ci);
  const float2 s2r = add2(br, dr), s2i = add2(bi, di), s3r = sub2(br, dr), s3i = sub2(bi, di);
  ar = add2(s0r, s2r); ai = add2(s0i, s2i);
  cr = sub2(s0r, s2r); ci = sub2(s0i, s2i);
  br = add2(s1r, s3i); bi = sub2(s1i, s3r);
}

// fft16x2 whose outputs 0, 1, 14 and 15 (slots perm16(k)) are not needed and not computed
AIP_HD void fft16x2_out_z2(float2 (&r)[16], float2 (&i)[16]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
    radix4x2(r[a], i[a], r[a + 4], i[a + 4], r[a + 8], i[a + 8], r[a + 12], i[a + 12]);
  // slot group q holds bins q, q + 4, q + 8, q + 12 in positions A, B, C, D
  radix4x2_noA(r[0], i[0], r[1], i[1], r[2], i[2], r[3], i[3]);          // bin 0 dropped
  radix4x2_w1<1>(r[4], i[4], r[5], i[5], r[6], i[6], r[7], i[7]);        // bin 1 dropped
  radix4x2_w2<2>(r[8], i[8], r[9], i[9], r[10], i[10], r[11], i[11]);    // bin 14 dropped
  radix4x2_w3<2>(r[12], i[12], r[13], i[13], r[14], i[14], r[15], i[15]);  // bin 15 dropped
}

// Same transform when inputs 0, 1, 14 and 15 are zero (never read: the slots may hold anything)
AIP_HD void fft16x2_in_z2(float2 (&r)[16], float2 (&i)[16]) {
  radix4x2_a0(r[0], i[0], r[4], i[4], r[8], i[8], r[12], i[12]);
  radix4x2_a0(r[1], i[1], r[5], i[5], r[9], i[9], r[13], i[13]);
  radix4x2_d0(r[2], i[2], r[6], i[6], r[10], i[10], r[14], i[14]);
  radix4x2_d0(r[3], i[3], r[7], i[7], r[11], i[11], r[15], i[15]);
  fft16x2_tail(r, i);
}

// Exchange buffer (both directions): float2 slot ((p*2 + c)*16 + n1)*33 + frame holds component c (0 = re,
// 1 = im) of the stage-1 outputs (n1, k2 = ja) in .x and (n1, k2 = jb) in .y, where (ja, jb) = (p, 16 - p) for
// pair-job p > 0 and (0, 8) for p = 0 -- i.e. exactly the packed operand a stage-2 thread feeds to fft16x2.
AIP_HD constexpr int exch_slot(int p, int c, int n1) { return ((p * 2 + c) * 16 + n1) * kXP; }
AIP_HD constexpr int job_a(int p) { return p; }
AIP_HD constexpr int job_b(int p) { return p == 0 ? 8 : 16 - p; }

// Per-lane constants of the stage that runs with lane = n1 (forward stage 1, inverse stage B): the 16
// inter-stage twiddles W256^(n1 k2) live in registers; the 32 window taps the lane touches come from a small
// shared-memory table (row n1, pitch 36 floats: four LDS.128 quarter-warp phases, conflict-free), because
// 64 packed data registers + 64 constants do not fit the 128-register budget of a 512-thread CTA.
struct LaneConst {
  float twr[16], twi[16];
};
constexpr int kWinPitch = 36;
constexpr int kWinTable = 16 * kWinPitch;      // floats

// tw_s[n1*17 + k2] = W256^(n1 k2): the per-lane twiddle rows, staged in shared memory once per CTA.  The lanes
// read their row from THERE, not from the __constant__ table: an indexed constant load serialises over the 16
// distinct lane addresses, and ptxas re-materialises constant loads inside the tile loop whenever the stage runs
// out of registers (measured: 9 such LDCs per tile cost 25 % of the forward kernel).
constexpr int kTwPitch = 17;                   // float2 per row (odd: conflict-free LDS.64 over n1)
constexpr int kTwTable = 16 * kTwPitch;        // float2
AIP_HD void twiddle_table_fill(float2* tw_s, int tid, int nthreads) {
  for (int idx = tid; idx < 256; idx += nthreads) {
    const int n1 = idx >> 4, k2 = idx & 15;
    tw_s[n1 * kTwPitch + k2] = kTw256[(n1 * k2) & 255];
  }
}

AIP_HD void lane_const_init(LaneConst& c, const float2* tw_s, int n1) {
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float2 t = tw_s[n1 * kTwPitch + k2];
    c.twr[k2] = t.x;
    c.twi[k2] = t.y;
  }
}

// win_s[n1*36 + 2*n2 + c] = window[2*(n1 + 16*n2) + c] * scale;  scale: 0.5 forward (split pass), 1/512
// inverse (irfft norm).  window: n_fft centre-padded taps.
AIP_HD void window_table_fill(float* win_s, const float* window, float scale, int tid, int nthreads) {
  for (int idx = tid; idx < 16 * 32; idx += nthreads) {
    const int n1 = idx >> 5, j = idx & 31;
    win_s[n1 * kWinPitch + j] = window[2 * (n1 + 16 * (j >> 1)) + (j & 1)] * scale;
  }
}

// ---------------------------------------------------------------------------------------------------
// Forward, stage 1.  One call = the (n1) column of TWO frames (fa in .x, fb in .y): 2 x 16 strided float2
// loads of the staged waveform, window, packed 16-point DFT over n2, inter-stage twiddle, 2 x 16 float2
// stores into the exchange buffer.
// ---------------------------------------------------------------------------------------------------
// ZP = number of leading AND trailing n2 groups (32 samples each) whose window taps are all zero -- librosa's
// pad_center of a win_length < n_fft window (config.py: win 384 in n_fft 512 => 64 zero taps each side => ZP = 2):
// those samples are neither loaded nor multiplied, and the first radix-4 pass runs its pruned form.
struct NoWait { AIP_HM void operator()() const {} };

// `before_store` runs between the DFT and the twiddle / store loop: the exchange buffer is not touched before it
template <int ZP, class Before>
AIP_HD void fwd_stage1(const float* tile, float2* exch, const float* win_s, int hop, int fa, int fb, int n1,
                       const LaneConst& c, Before& before_store) {
  float2 r[16], i[16];
  const float* sa = tile + fa * hop + 2 * n1;
  const float* sb = tile + fb * hop + 2 * n1;
  const float4* wrow = reinterpret_cast<const float4*>(win_s + n1 * kWinPitch);
#pragma unroll
  for (int j = ZP / 2; j < 8 - ZP / 2; ++j) {
    const float4 w = wrow[j];        // (we, wo) of n2 = 2j and 2j + 1
    const float2 za0 = *reinterpret_cast<const float2*>(sa + 64 * j);
    const float2 zb0 = *reinterpret_cast<const float2*>(sb + 64 * j);
    const float2 za1 = *reinterpret_cast<const float2*>(sa + 64 * j + 32);
    const float2 zb1 = *reinterpret_cast<const float2*>(sb + 64 * j + 32);
    r[2 * j] = make_float2(za0.x * w.x, zb0.x * w.x);
    i[2 * j] = make_float2(za0.y * w.y, zb0.y * w.y);
    r[2 * j + 1] = make_float2(za1.x * w.z, zb1.x * w.z);
    i[2 * j + 1] = make_float2(za1.y * w.w, zb1.y * w.w);
  }
  if (ZP == 2) fft16x2_in_z2(r, i);
  else fft16x2(r, i);
  before_store();
  float2* da = exch + n1 * kXP + fa;
  float2* db = exch + n1 * kXP + fb;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int ka = job_a(p), kb = job_b(p);
    // frame fa (.x lanes) and frame fb (.y lanes); the twiddled results are written as (job a, job b) pairs
    float ar_a = r[perm16(ka)].x, ai_a = i[perm16(ka)].x, br_a = r[perm16(kb)].x, bi_a = i[perm16(kb)].x;
    float ar_b = r[perm16(ka)].y, ai_b = i[perm16(ka)].y, br_b = r[perm16(kb)].y, bi_b = i[perm16(kb)].y;
    if (ka > 0) { cmul(ar_a, ai_a, c.twr[ka], c.twi[ka]); cmul(ar_b, ai_b, c.twr[ka], c.twi[ka]); }
    cmul(br_a, bi_a, c.twr[kb], c.twi[kb]);
    cmul(br_b, bi_b, c.twr[kb], c.twi[kb]);
    da[exch_slot(p, 0, 0)] = make_float2(ar_a, br_a);
    da[exch_slot(p, 1, 0)] = make_float2(ai_a, bi_a);
    db[exch_slot(p, 0, 0)] = make_float2(ar_b, br_b);
    db[exch_slot(p, 1, 0)] = make_float2(ai_b, bi_b);
  }
}

// Split-pass twiddles of one pair-job, register resident (a stage-2 / stage-A warp keeps its job p for
// the whole kernel), stored the way the PACKED split pass consumes them.
//   p > 0:  (wr2[j], wi2[j]) = ( (Re W^k, Re W^k'), (Im W^k, -Im W^k') ),  W = W512, k = p + 16 j, k' = p + 16 (15 - j)
//   p = 0:  .x = W512^(16 (j + 1)) (job 0, pair j + 1 <-> 15 - j, j = 0..6),  .y = W512^(8 + 16 j) (job 8, pair j <-> 15 - j)
struct PairTw {
  float2 wr2[8], wi2[8];
};

AIP_HD void pair_tw_init(PairTw& w, int p) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (p != 0) {
      const float2 a = kTw512[p + 16 * j], b = kTw512[p + 16 * (15 - j)];
      w.wr2[j] = make_float2(a.x, b.x);
      w.wi2[j] = make_float2(a.y, -b.y);
    } else {
      const float2 a = kTw512[16 * (j + 1)], b = kTw512[8 + 16 * j];
      w.wr2[j] = make_float2(a.x, b.x);
      w.wi2[j] = make_float2(a.y, b.y);
    }
  }
}
// scalar view for p > 0: W512^(p + 16 k1), k1 = 0..15
AIP_HD float pair_tw_r(const PairTw& w, int k1) { return k1 < 8 ? w.wr2[k1].x : w.wr2[15 - k1].y; }
AIP_HD float pair_tw_i(const PairTw& w, int k1) { return k1 < 8 ? w.wi2[k1].x : -w.wi2[15 - k1].y; }

// Split pass for one (k, 256-k) pair, scalar: Zk = Zc[k], Zn = Zc[256-k], (wr, wi) = W512^k.  Both results go to
// the emitter as one packed pair (.x = bin k at offset o_lo, .y = bin 256-k at offset o_hi) so that the magnitude
// epilogue runs on FP32x2.
template <class Emit>
AIP_HD void fwd_pair(float zkr, float zki, float znr, float zni, float wr, float wi, typename Emit::off_t o_lo,
                     typename Emit::off_t o_hi, Emit& emit) {
  const float er = zkr + znr, ei = zki - zni;
  const float orr = zkr - znr, oi = zki + zni;
  const float tr = orr * wr - oi * wi;
  const float ti = orr * wi + oi * wr;
  emit.template put2<1, -1>(o_lo, o_hi, make_float2(er + ti, er - ti), make_float2(ei - tr, ei + tr));
}

// Split pass for TWO pairs at once, packed.  After fft16x2 a stage-2 thread of pair-job p > 0 holds, in slot j,
// U = (Za[j], Zb[j]) with Za[j] = Zc[p + 16 j] and Zb[j] = Zc[(16 - p) + 16 j].  For k = p + 16 j and
// k' = p + 16 (15 - j) the partners are Zc[256 - k] = Zb[15 - j] and Zc[256 - k'] = Zb[j], i.e.
//     U = slot j      = (Zk , Zn')          V = slot 15 - j = (Zk', Zn)
// so (U, swap(V)) are the (Zk, Zn) operands of pair k in lane .x and, with the roles of Z and its partner
// exchanged, of pair k' in lane .y.  Exchanged roles flip the sign of Re O and Im E in lane .y; with the lane-.y
// twiddle conjugated (PairTw) every product comes out right up to the sign of the IMAGINARY outputs of lane .y,
// which the emitter fixes only when it actually stores imaginary parts or phases.
template <class Emit>
AIP_HD void fwd_pair2(float2 ur, float2 ui, float2 vr, float2 vi, float2 wr2, float2 wi2, typename Emit::off_t lo_x,
                      typename Emit::off_t lo_y, typename Emit::off_t hi_x, typename Emit::off_t hi_y, Emit& emit) {
  const float2 vsr = swap2(vr), vsi = swap2(vi);
  const float2 er = add2(ur, vsr), oi = add2(ui, vsi);       // (Re E, Re E'), (Im O, Im O')
  const float2 dr = sub2(ur, vsr), di = sub2(ui, vsi);       // (Re O, -Re O'), (Im E, -Im E')
  const float2 tr = fma2(neg2(oi), wi2, mul2(dr, wr2));      // (Re T, -Re T'),  T = W O
  const float2 ti = fma2(dr, wi2, mul2(oi, wr2));            // (Im T,  Im T')
  emit.template put2<1, -1>(lo_x, lo_y, add2(er, ti), sub2(di, tr));     // X[k], X[k']:   (er + ti, ei - tr)
  emit.template put2<-1, 1>(hi_x, hi_y, sub2(er, ti), add2(di, tr));     // X[256-k], X[256-k']: (er - ti, -(ei + tr))
}

// ---------------------------------------------------------------------------------------------------
// Forward, stage 2.  One (frame f, pair-job p) job, p = 0..7, in two steps so that the exchange buffer
// can be handed back to the stage-1 warps as soon as it has been read:
//   fwd_stage2_load     32 float2 of jobs (p, 16-p) or (0, 8) -> registers
//   fwd_stage2_compute  two 16-point DFTs over n1, the split pass, 32 (p>0) / 33 (p=0) bins to the emitter
// ---------------------------------------------------------------------------------------------------
AIP_HD void fwd_stage2_load(const float2* exch, int f, int p, float2 (&r)[16], float2 (&i)[16]) {
  const float2* sr = exch + exch_slot(p, 0, 0) + f;
  const float2* si = exch + exch_slot(p, 1, 0) + f;
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    r[n1] = sr[n1 * kXP];
    i[n1] = si[n1 * kXP];
  }
}

// Two pairs at once with the SAME roles in both lanes (pair-job 0): lane .x is a pair of job 0, lane .y one of job 8.
template <class Emit>
AIP_HD void fwd_pair2s(float2 zkr, float2 zki, float2 znr, float2 zni, float2 wr2, float2 wi2, typename Emit::off_t lo_x,
                       typename Emit::off_t lo_y, typename Emit::off_t hi_x, typename Emit::off_t hi_y, Emit& emit) {
  const float2 er = add2(zkr, znr), ei = sub2(zki, zni);
  const float2 orr = sub2(zkr, znr), oi = add2(zki, zni);
  const float2 tr = fma2(neg2(oi), wi2, mul2(orr, wr2));
  const float2 ti = fma2(orr, wi2, mul2(oi, wr2));
  emit.template put2<1, 1>(lo_x, lo_y, add2(er, ti), sub2(ei, tr));
  emit.template put2<-1, -1>(hi_x, hi_y, sub2(er, ti), add2(ei, tr));
}

struct NoHook { AIP_HM void operator()() const {} };

// `hook` runs once: between the two DFTs and the split pass (kHookAt < 0) or after kHookAt + 1 of the 8 split-pass / epilogue
// rounds (see fwd_phase2: the late release of the exchange buffer)
template <class Emit, class Hook, int kHookAt = -1>
AIP_HD void fwd_stage2_compute(float2 (&r)[16], float2 (&i)[16], const PairTw& w, int p, Emit& emit, Hook& hook) {
  fft16x2(r, i);     // .x = job a, .y = job b
  if (kHookAt < 0) hook();
  if (p != 0) {
    emit.rows(p, 256 - p);
#if defined(AIP_SPLIT_SCALAR)
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1)
      fwd_pair(r[perm16(k1)].x, i[perm16(k1)].x, r[perm16(15 - k1)].y, i[perm16(15 - k1)].y, pair_tw_r(w, k1),
               pair_tw_i(w, k1), emit.lo(k1), emit.hi(k1), emit);
#else
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      fwd_pair2(r[perm16(j)], i[perm16(j)], r[perm16(15 - j)], i[perm16(15 - j)], w.wr2[j], w.wi2[j],
                emit.lo(j), emit.lo(15 - j), emit.hi(j), emit.hi(15 - j), emit);
      if (j == kHookAt) hook();
    }
#endif
  } else {
    // Job 0 (.x: bins 16 k1, pairs k1 <-> 16 - k1) and job 8 (.y: bins 8 + 16 k1, pairs k1 <-> 15 - k1).  Pair j + 1 of
    // job 0 and pair j of job 8 share their partner slot 15 - j; their own operands sit in slots j + 1 (.x) and j (.y),
    // so one register move per component lines them up for the packed split pass.
    const float z0r = r[perm16(0)].x, z0i = i[perm16(0)].x;
    emit.template put2<1, 1>(emit.bin(0), emit.bin(256), make_float2(2.0f * (z0r + z0i), 2.0f * (z0r - z0i)),
                             make_float2(0.0f, 0.0f));
    emit.put1(emit.bin(128), 2.0f * r[perm16(8)].x, -2.0f * i[perm16(8)].x);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float2 zkr = make_float2(r[perm16(j + 1)].x, r[perm16(j)].y);
      const float2 zki = make_float2(i[perm16(j + 1)].x, i[perm16(j)].y);
      fwd_pair2s(zkr, zki, r[perm16(15 - j)], i[perm16(15 - j)], w.wr2[j], w.wi2[j], emit.bin(16 * (j + 1)),
                 emit.bin(8 + 16 * j), emit.bin(256 - 16 * (j + 1)), emit.bin(248 - 16 * j), emit);
      if (j == kHookAt) hook();
    }
    if (kHookAt >= 7) hook();
    fwd_pair(r[perm16(7)].y, i[perm16(7)].y, r[perm16(8)].y, i[perm16(8)].y, w.wr2[7].y, w.wi2[7].y, emit.bin(120),
             emit.bin(136), emit);
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
  float2 r[16], i[16];          // .x = job a, .y = job b
  if (live) {
    if (p != 0) {
      load.rows(p, 256 - p);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        float xr, xi, yr, yi;
        load.lo(k1, xr, xi);
        load.hi(k1, yr, yi);
        inv_pair(xr, xi, yr, yi, pair_tw_r(w, k1), pair_tw_i(w, k1), r[k1].x, i[k1].x, r[15 - k1].y, i[15 - k1].y);
      }
    } else {
      float xr, xi, yr, yi;
      load.rows(0, 256);
      load.lo(0, xr, xi);
      load.hi(0, yr, yi);
      r[0].x = xr + yr; i[0].x = xr - yr;
#pragma unroll
      for (int k1 = 1; k1 < 8; ++k1) {
        load.lo(k1, xr, xi);
        load.hi(k1, yr, yi);
        inv_pair(xr, xi, yr, yi, w.wr2[k1 - 1].x, w.wi2[k1 - 1].x, r[k1].x, i[k1].x, r[16 - k1].x, i[16 - k1].x);
      }
      load.lo(8, xr, xi);
      r[8].x = 2.0f * xr; i[8].x = -2.0f * xi;
      load.rows(8, 248);
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) {
        load.lo(k1, xr, xi);
        load.hi(k1, yr, yi);
        inv_pair(xr, xi, yr, yi, w.wr2[k1].y, w.wi2[k1].y, r[k1].y, i[k1].y, r[15 - k1].y, i[15 - k1].y);
      }
    }
    load.done();        // every bin of this job is in registers (a staged loader refills its buffer from here)
    fft16x2(i, r);      // inverse transform: swapped roles
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) { r[q] = make_float2(0.0f, 0.0f); i[q] = make_float2(0.0f, 0.0f); }
  }
  before_store();      // the exchange buffer must have been released by the previous tile's readers
  float2* dr = exch + exch_slot(p, 0, 0) + f;
  float2* di = exch + exch_slot(p, 1, 0) + f;
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    dr[n1 * kXP] = r[perm16(n1)];
    di[n1 * kXP] = i[perm16(n1)];
  }
}

// ---------------------------------------------------------------------------------------------------
// Inverse, stage B.  One call = the (n1) column of TWO frames (fa in .x, fb in .y): conj inter-stage twiddle,
// packed inverse 16-point DFT over k2, synthesis window (carrying the 1/512 of irfft), windowed samples
// written IN PLACE into the exchange buffer, which thereby becomes the frame buffer: float2 slot
// (m*33 + f) = samples (2m, 2m+1) of frame f, m = n1 + 16 n2 (a thread reads and writes the same 2 x 16 slots).
// ---------------------------------------------------------------------------------------------------
// ZP = 2: the first / last 64 synthesis-window taps are zero (win 384 in n_fft 512): those samples are neither computed
// nor stored -- their frame-buffer slots keep stale data, so only an overlap-add that skips them may follow
// (inv_ola_fast<192, 0, 1>).
template <int ZP>
AIP_HD void inv_stageB(float2* exch, const float* win_s, int fa, int fb, int n1, const LaneConst& c) {
  float2 r[16], i[16];
  float2* pa = exch + n1 * kXP + fa;
  float2* pb = exch + n1 * kXP + fb;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int ka = job_a(p), kb = job_b(p);
    const float2 ra = pa[exch_slot(p, 0, 0)], ia = pa[exch_slot(p, 1, 0)];   // frame fa: (job a, job b)
    const float2 rb = pb[exch_slot(p, 0, 0)], ib = pb[exch_slot(p, 1, 0)];   // frame fb
    float ar_a = ra.x, ai_a = ia.x, br_a = ra.y, bi_a = ia.y;
    float ar_b = rb.x, ai_b = ib.x, br_b = rb.y, bi_b = ib.y;
    if (ka > 0) { cmul(ar_a, ai_a, c.twr[ka], -c.twi[ka]); cmul(ar_b, ai_b, c.twr[ka], -c.twi[ka]); }
    cmul(br_a, bi_a, c.twr[kb], -c.twi[kb]);
    cmul(br_b, bi_b, c.twr[kb], -c.twi[kb]);
    r[ka] = make_float2(ar_a, ar_b); i[ka] = make_float2(ai_a, ai_b);
    r[kb] = make_float2(br_a, br_b); i[kb] = make_float2(bi_a, bi_b);
  }
  if (ZP == 2) fft16x2_out_z2(i, r);
  else fft16x2(i, r);
  const float4* wrow = reinterpret_cast<const float4*>(win_s + n1 * kWinPitch);
#pragma unroll
  for (int j = ZP / 2; j < 8 - ZP / 2; ++j) {
    const float4 w = wrow[j];
    const float2 zr0 = r[perm16(2 * j)], zi0 = i[perm16(2 * j)];
    const float2 zr1 = r[perm16(2 * j + 1)], zi1 = i[perm16(2 * j + 1)];
    pa[(2 * j) * 16 * kXP] = make_float2(zr0.x * w.x, zi0.x * w.y);
    pb[(2 * j) * 16 * kXP] = make_float2(zr0.y * w.x, zi0.y * w.y);
    pa[(2 * j + 1) * 16 * kXP] = make_float2(zr1.x * w.z, zi1.x * w.w);
    pb[(2 * j + 1) * 16 * kXP] = make_float2(zr1.y * w.z, zi1.y * w.w);
  }
}

// clip-edge tiles after inv_stageB<2>: the general overlap-add reads every tap, so the skipped ones must read as zero
AIP_HD void inv_stageB_zero_pruned(float2* exch, int fa, int fb, int n1) {
  float2* pa = exch + n1 * kXP + fa;
  float2* pb = exch + n1 * kXP + fb;
  const float2 z = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2)
    if (n2 < 2 || n2 >= 14) { pa[n2 * 16 * kXP] = z; pb[n2 * 16 * kXP] = z; }
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
