// Inverse transform + Griffin-Lim: sm_100a kernels + their launchers + the back-end entry points of the C ABI.
//
//   istft512_kernel      persistent; tile = 32 frames -> FO hops of output; split-pass prologue,
//                        register inverse FFT, synthesis window, overlap-add and window-sum-square
//                        normalisation out of shared memory (no global atomics, halo frames recomputed),
//                        optional per-clip peak for the fused normalisation.
//   istft512_tma_kernel  the same with stage A's rows staged by 4-D TMA tensor boxes (tunable AIP_INV_TMA=1).
//   gl_update*_kernel    the Griffin-Lim phase update between the two transforms (unfused fallbacks).
//   istft_generic_*      any power-of-two n_fft in [32, 4096] (or odd hop).
#include "aip_device.cuh"
#include "aip_host.h"

namespace aip {

// Stage A reads, per pair-job p, the 32 (p = 0: 33) rows p + 16 j and 256 - p - 16 j of the tile's 32 frames: one 128- or
// 256-byte segment per row and array.  With direct loads the stage waits for HBM on every row (ncu long_scoreboard 3-8 stall
// cycles per issued instruction).  Here the warp asks the L2 for the NEXT tile's segments while it works on this one -- lane j
// takes row j of the job, so it costs one or two prefetch instructions per thread, array and tile, not one per load.
template <int kElemBytes>
__device__ __forceinline__ void inv_prefetch_rows(const void* array, const InvParams& P, const TileCursor& c, int p, int lane) {
  int t0 = c.tt * P.g.FO - P.g.HL;
  if (t0 < 0) t0 = 0;
  if (t0 >= P.n_frames) return;
  int row;
  if (p != 0) row = lane < 16 ? p + 16 * lane : 256 - p - 16 * (lane - 16);
  else row = lane < 17 ? 16 * lane : 8 + 16 * (lane - 17);           // rows 16 j (17 of them) and 8 + 16 j, j = 0 .. 14
  const char* q = static_cast<const char*>(array) + (((long long)c.b * kBins + row) * P.T + t0) * kElemBytes;
  prefetch_l2(q);
  prefetch_l2(q + 128);                                               // 32 frames x 4 bytes may straddle two lines, x 8 bytes span two
  if (kElemBytes == 8) prefetch_l2(q + 256);
  if (p == 0 && lane == 0) {                                          // the 33rd row of job (0, 8): bin 248
    const char* r = static_cast<const char*>(array) + (((long long)c.b * kBins + 248) * P.T + t0) * kElemBytes;
    prefetch_l2(r); prefetch_l2(r + 128);
    if (kElemBytes == 8) prefetch_l2(r + 256);
  }
}

// The arrays a mode's loader reads.  Measured (1024 x 10 s / 2048 x 5 s, hop 192): magnitude + phase input 0.817 -> 0.710 ms and
// 0.840 -> 0.736 ms, the TMA-staged Griffin-Lim inverse (its |S| loads) 51.6 -> 47.2 ms per 32 iterations; for the directly
// loaded COMPLEX inputs it does not pay (0.554 -> 0.549 ms at T = 834, 0.580 -> 0.589 ms at T = 417; Griffin-Lim with three
// directly loaded arrays 57.8 -> 61.2 ms): those modes issue no requests.
template <int kMode>
__device__ __forceinline__ void inv_prefetch_tile(const InvParams& P, const TileCursor& c, int p, int lane) {
  if (kMode == INV_SPEC || kMode == INV_GL) {
    return;
  } else {
    inv_prefetch_rows<4>(P.mag, P, c, p, lane);
    if (P.phase) inv_prefetch_rows<4>(P.phase, P, c, p, lane);
    if (P.blend_in) { inv_prefetch_rows<4>(P.blend_in, P, c, p, lane); inv_prefetch_rows<4>(P.blend_mask, P, c, p, lane); }
  }
}

struct WaitBefore {
  uint64_t* bar;
  uint32_t parity;
  bool enabled;
  __device__ __forceinline__ void operator()() const { if (enabled) mbar_wait(bar, parity); }
};

// Warp-specialised, persistent, one CTA per SM (same plumbing as the forward kernel):
//   stage-A warps (lane = frame): HBM -> registers, split-pass prologue, 2 x inverse 16-pt DFT -> exch[i % 3]
//   stage-B warps (lane = n1): twiddle, inverse 16-pt DFT, synthesis window (in place: the exchange buffer
//   becomes the frame buffer) -> named barrier -> overlap-add + 1/window-sum-square -> HBM -> release
// Stage A computes tile i+1 in registers while stage B works on tile i and only then waits for an exchange
// buffer; a ring of n_bufs buffers is supported, two measured fastest (see kInvBufsDefault).
template <int kMode, int kFast>
__global__ void __launch_bounds__(kFwdThreads, 1) istft512_kernel(const InvParams P) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2 * kInvBufs];
  __shared__ __align__(8) float wtab_s[kMaxWtab];
  __shared__ __align__(16) float win_s[kWinTable];
  __shared__ __align__(8) float2 tw_s[kTwTable];
  window_table_fill(win_s, P.window, 1.0f / 512.0f, threadIdx.x, blockDim.x);
  twiddle_table_fill(tw_s, threadIdx.x, blockDim.x);
  const float* wtab = (P.wss_ref >= 0 && P.hop <= kMaxWtab) ? wtab_s : nullptr;
  if (wtab)
    for (int r = threadIdx.x; r < P.hop; r += blockDim.x) wtab_s[r] = P.inv_wss[P.wss_ref + r];
  uint64_t* exch_full = bars;                // [3] count 8 (stage-A warps)
  uint64_t* exch_empty = bars + kInvBufs;    // [3] count 8 (stage-B warps)
  float2* exch0 = reinterpret_cast<float2*>(smem);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kInvBufs; ++i) {
      mbar_init(exch_full + i, kThreads / 32);
      mbar_init(exch_empty + i, kThreads / 32);
    }
  }
  __syncthreads();
  const int first = blockIdx.x * P.tiles_per_cta;
  int n = P.n_tiles - first;
  if (n > P.tiles_per_cta) n = P.tiles_per_cta;
  if (n <= 0) return;
  TileCursor c = tile_cursor(first, P.tiles_per_clip);
  if (tid < kThreads) {
    PairTw w;
    pair_tw_init(w, tid >> 5);
    int es = 0, use = 0;                      // ring slot and how often it has been used before
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      if (P.l2_prefetch && i + 1 < n) {
        TileCursor cn = c;
        tile_advance(cn, P.tiles_per_clip);
        inv_prefetch_tile<kMode>(P, cn, tid >> 5, tid & 31);
      }
      WaitBefore wb{exch_empty + es, (uint32_t)((use - 1) & 1), use >= 1};
      inv_phase0<kMode>(P, tid, c, exch0 + es * kExch, w, wb);
      mbar_arrive_warp(exch_full + es);
      tile_advance(c, P.tiles_per_clip);
      if (++es == P.n_bufs) { es = 0; ++use; }
    }
  } else {
    const int btid = tid - kThreads;
    LaneConst lc;
    lane_const_init(lc, tw_s, btid & 15);
    int es = 0, use = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      float2* exch = exch0 + es * kExch;
      mbar_wait(exch_full + es, (uint32_t)(use & 1));
      inv_phase1<kFast>(P, btid, c, exch, win_s, lc);
      named_bar_sync(1, kThreads);
      inv_phase2<kFast>(P, btid, c, exch, wtab);
      mbar_arrive_warp(exch_empty + es);
      tile_advance(c, P.tiles_per_clip);
      if (++es == P.n_bufs) { es = 0; ++use; }
    }
  }
}

// ---- complex input staged by TMA tensor tiles ---------------------------------------------------------
// The direct-load kernel above keeps at most a few spectrum loads per thread in flight (the 64 data registers of the
// packed codelet leave no room for more), which bounds it by HBM latency.  Here every stage-A warp owns a private
// staging slot that TMA fills one tile ahead: the spectrogram is described to the TMA unit as the
// 4-D tensor (2t, p, j, b) with bin k = 16 j + p, so the 16 rows p + 16 j a pair-job reads are ONE box of
// 68 floats x 1 x 16 x 1, the 16 partner rows 256 - p - 16 j a second one, and frames outside [0, n_frames) arrive
// as zeros.  Needs an even T (row pitch 8 T bytes must be a multiple of 16) and a 16-byte aligned base.
// The TMA unit needs the innermost start coordinate 16-byte aligned (an odd first frame raises an illegal-instruction
// fault, tools/microbench/tma3d_probe.cu), so a box starts at the even frame below t0 and is 34 frames wide.
// ONE slot per warp (8 x 9 KB): the warp moves its 32 / 33 bins into registers, asks for the NEXT tile's boxes right away
// (InvLoadStaged::done) and only then runs its codelets and waits for an exchange buffer -- a whole stage-A phase hides the
// copy -- which leaves room for the ring of TWO exchange buffers the direct-load kernel has (round 1 staged two slots per
// warp, 142 KB, next to a single exchange buffer and lost to the direct loads: 0.568 against 0.527 ms).
constexpr int kStageFr = kFR + 2;                       // frames per staged row
constexpr int kStageB = 592;                            // float2 offset of region B (128-byte aligned: 4736 B)
constexpr int kStageSlot = 1136;                        // float2 per slot (9088 B): A = rows 0..16 (17 x 34), B = 16 rows
constexpr int kStageBytesWarp = kStageSlot * 8;         // one slot

// one elected lane: the two (p > 0) or three (p = 0) boxes of a tile into a staging slot
__device__ __forceinline__ void inv_issue_stage(const InvParams& P, const CUtensorMap* map16, const CUtensorMap* map1,
                                                const TileCursor& c, int p, float2* slot, uint64_t* bar) {
  const int t0e = (c.tt * P.g.FO - P.g.HL) & ~1;
  if (p != 0) {
    mbar_expect_tx(bar, 2u * 16u * kStageFr * 8u);
    tma_load_4d(slot, map16, 2 * t0e, p, 0, c.b, bar);
    tma_load_4d(slot + kStageB, map16, 2 * t0e, 16 - p, 0, c.b, bar);
  } else {
    mbar_expect_tx(bar, 33u * kStageFr * 8u);
    tma_load_4d(slot, map16, 2 * t0e, 0, 0, c.b, bar);
    tma_load_4d(slot + 16 * kStageFr, map1, 2 * t0e, 0, 16, c.b, bar);
    tma_load_4d(slot + kStageB, map16, 2 * t0e, 8, 0, c.b, bar);
  }
}

struct InvLoadStaged {      // stage-A loader out of the warp's staging slot (lane = frame)
  const float2* slot;       // + lane + (t0 & 1)
  const float2* plo;
  const float2* phi;
  // refill: the next tile's boxes into the same slot as soon as this tile's bins are in registers
  const InvParams& P;
  const CUtensorMap* map16;
  const CUtensorMap* map1;
  const TileCursor& next;
  int p;
  float2* slot_base;
  uint64_t* full;
  bool refill;
  __device__ __forceinline__ void rows(int k_lo, int k_hi) {
    if (k_hi == 256) { plo = slot; phi = slot + 16 * kStageFr; }                               // job 0: rows j and 16 - j of region A
    else if (k_lo == 8) { plo = slot + kStageB; phi = slot + kStageB + 15 * kStageFr; }       // job 8: region B
    else { plo = slot; phi = slot + kStageB + 15 * kStageFr; }                                 // p > 0: A = p + 16 j, B = (16-p) + 16 j'
  }
  __device__ __forceinline__ void lo(int j, float& xr, float& xi) const { const float2 v = plo[j * kStageFr]; xr = v.x; xi = v.y; }
  __device__ __forceinline__ void hi(int j, float& xr, float& xi) const { const float2 v = phi[-j * kStageFr]; xr = v.x; xi = v.y; }
  __device__ __forceinline__ void done() const {
    __syncwarp();
    if (refill && (threadIdx.x & 31) == 0) {
      fence_proxy_async();
      inv_issue_stage(P, map16, map1, next, p, slot_base, full);
    }
  }
};

template <int kFast>
__global__ void __launch_bounds__(kFwdThreads, 1) istft512_tma_kernel(const InvParams P, const __grid_constant__ CUtensorMap map16,
                                                                       const __grid_constant__ CUtensorMap map1) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2 * kInvBufs + 8];
  __shared__ __align__(8) float wtab_s[kMaxWtab];
  __shared__ __align__(16) float win_s[kWinTable];
  __shared__ __align__(8) float2 tw_s[kTwTable];
  window_table_fill(win_s, P.window, 1.0f / 512.0f, threadIdx.x, blockDim.x);
  twiddle_table_fill(tw_s, threadIdx.x, blockDim.x);
  const float* wtab = (P.wss_ref >= 0 && P.hop <= kMaxWtab) ? wtab_s : nullptr;
  if (wtab)
    for (int r = threadIdx.x; r < P.hop; r += blockDim.x) wtab_s[r] = P.inv_wss[P.wss_ref + r];
  uint64_t* exch_full = bars;
  uint64_t* exch_empty = bars + kInvBufs;
  uint64_t* stage_full = bars + 2 * kInvBufs;      // [warp], count 1 (+ tx bytes)
  float2* stage0 = reinterpret_cast<float2*>(smem);                         // 8 warps x 1 slot
  float2* exch0 = stage0 + 8 * kStageSlot;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kInvBufs; ++i) {
      mbar_init(exch_full + i, kThreads / 32);
      mbar_init(exch_empty + i, kThreads / 32);
    }
    for (int i = 0; i < 8; ++i) mbar_init(stage_full + i, 1);
  }
  __syncthreads();
  const int first = blockIdx.x * P.tiles_per_cta;
  int n = P.n_tiles - first;
  if (n > P.tiles_per_cta) n = P.tiles_per_cta;
  if (n <= 0) return;
  TileCursor c = tile_cursor(first, P.tiles_per_clip);
  if (tid < kThreads) {
    const int warp = tid >> 5, lane = tid & 31;
    PairTw w;
    pair_tw_init(w, warp);
    float2* my_stage = stage0 + warp * kStageSlot;
    uint64_t* my_full = stage_full + warp;
    TileCursor c1 = c;                                  // cursor of the next tile (the refill)
    if (lane == 0) inv_issue_stage(P, &map16, &map1, c1, warp, my_stage, my_full);
    tile_advance(c1, P.tiles_per_clip);
    int es = 0, use = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      mbar_wait(my_full, (uint32_t)(i & 1));
      InvLoadStaged load{my_stage + lane + ((c.tt * P.g.FO - P.g.HL) & 1), nullptr, nullptr,
                         P, &map16, &map1, c1, warp, my_stage, my_full, i + 1 < n};
      WaitBefore wb{exch_empty + es, (uint32_t)((use - 1) & 1), use >= 1};
      inv_stageA(exch0 + es * kExch, w, lane, warp, true, load, wb);
      mbar_arrive_warp(exch_full + es);
      tile_advance(c, P.tiles_per_clip);
      tile_advance(c1, P.tiles_per_clip);
      if (++es == P.n_bufs) { es = 0; ++use; }
    }
  } else {
    const int btid = tid - kThreads;
    LaneConst lc;
    lane_const_init(lc, tw_s, btid & 15);
    int es = 0, use = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      float2* exch = exch0 + es * kExch;
      mbar_wait(exch_full + es, (uint32_t)(use & 1));
      inv_phase1<kFast>(P, btid, c, exch, win_s, lc);
      named_bar_sync(1, kThreads);
      inv_phase2<kFast>(P, btid, c, exch, wtab);
      mbar_arrive_warp(exch_empty + es);
      tile_advance(c, P.tiles_per_clip);
      if (++es == P.n_bufs) { es = 0; ++use; }
    }
  }
}

// ---- Griffin-Lim inverse (INV_GL) with the two complex arrays staged by TMA -----------------------------------------
// The fused phase update reads THREE arrays per bin (rebuilt[it], rebuilt[it - 1], |S|); with direct loads stage A is bound by
// load latency (ncu, round 1: long_scoreboard 7.8 stall cycles per issued instruction, 36 B of spills).  Here the two complex
// arrays arrive through the same per-warp staging slot as above (one slot holds both: 2 x 9 KB per warp), |S| is read with
// one coalesced 4-byte load per bin, and a single exchange buffer fits beside the staging (145 + 66 KB).
struct InvLoadStagedGL {
  const float2* reb;        // staging slot of rebuilt[it] + lane + (t0 & 1); rebuilt[it - 1] sits kStageSlot float2 further
  const float2* plo;
  const float2* phi;
  const float* mag;         // |S| column: array + b F T + t
  int T;
  float alpha;
  int mlo, mhi;             // element offsets of the two row cursors in |S|
  const InvParams& P;
  const CUtensorMap* maps;  // [4]: rebuilt 16-row / 1-row boxes, previous 16-row / 1-row boxes
  const TileCursor& next;
  int p;
  float2* slot_base;
  uint64_t* full;
  bool refill;
  __device__ __forceinline__ void rows(int k_lo, int k_hi) {
    if (k_hi == 256) { plo = reb; phi = reb + 16 * kStageFr; }
    else if (k_lo == 8) { plo = reb + kStageB; phi = reb + kStageB + 15 * kStageFr; }
    else { plo = reb; phi = reb + kStageB + 15 * kStageFr; }
    mlo = k_lo * T; mhi = k_hi * T;
  }
  __device__ __forceinline__ void project(const float2* q, float m, float& xr, float& xi) const {
    const float2 r = q[0], t = q[kStageSlot];
    const float ax = r.x - alpha * t.x, ay = r.y - alpha * t.y;
    const float sc = fast_div(1.0f, fast_sqrt(ax * ax + ay * ay) + kFltMin);
    xr = (ax * sc) * m; xi = (ay * sc) * m;
  }
  __device__ __forceinline__ void lo(int j, float& xr, float& xi) const { project(plo + j * kStageFr, __ldcg(mag + mlo + j * 16 * T), xr, xi); }
  __device__ __forceinline__ void hi(int j, float& xr, float& xi) const { project(phi - j * kStageFr, __ldcg(mag + mhi - j * 16 * T), xr, xi); }
  __device__ __forceinline__ void done() const;
};

__device__ __forceinline__ void inv_issue_stage_gl(const InvParams& P, const CUtensorMap* maps, const TileCursor& c, int p,
                                                   float2* slot, uint64_t* bar) {
  const int t0e = (c.tt * P.g.FO - P.g.HL) & ~1;
  mbar_expect_tx(bar, 2u * (p != 0 ? 32u : 33u) * kStageFr * 8u);
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    float2* s = slot + a * kStageSlot;
    const CUtensorMap* m16 = maps + 2 * a;
    const CUtensorMap* m1 = maps + 2 * a + 1;
    if (p != 0) {
      tma_load_4d(s, m16, 2 * t0e, p, 0, c.b, bar);
      tma_load_4d(s + kStageB, m16, 2 * t0e, 16 - p, 0, c.b, bar);
    } else {
      tma_load_4d(s, m16, 2 * t0e, 0, 0, c.b, bar);
      tma_load_4d(s + 16 * kStageFr, m1, 2 * t0e, 0, 16, c.b, bar);
      tma_load_4d(s + kStageB, m16, 2 * t0e, 8, 0, c.b, bar);
    }
  }
}

__device__ __forceinline__ void InvLoadStagedGL::done() const {
  __syncwarp();
  if (refill && (threadIdx.x & 31) == 0) {
    fence_proxy_async();
    inv_issue_stage_gl(P, maps, next, p, slot_base, full);
  }
}

struct GlMaps { CUtensorMap m[4]; };

template <int kFast>
__global__ void __launch_bounds__(kFwdThreads, 1) istft512_gl_tma_kernel(const InvParams P, const __grid_constant__ GlMaps maps) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2 * kInvBufs + 8];
  __shared__ __align__(8) float wtab_s[kMaxWtab];
  __shared__ __align__(16) float win_s[kWinTable];
  __shared__ __align__(8) float2 tw_s[kTwTable];
  window_table_fill(win_s, P.window, 1.0f / 512.0f, threadIdx.x, blockDim.x);
  twiddle_table_fill(tw_s, threadIdx.x, blockDim.x);
  const float* wtab = (P.wss_ref >= 0 && P.hop <= kMaxWtab) ? wtab_s : nullptr;
  if (wtab)
    for (int r = threadIdx.x; r < P.hop; r += blockDim.x) wtab_s[r] = P.inv_wss[P.wss_ref + r];
  uint64_t* exch_full = bars;
  uint64_t* exch_empty = bars + kInvBufs;
  uint64_t* stage_full = bars + 2 * kInvBufs;
  float2* stage0 = reinterpret_cast<float2*>(smem);                         // 8 warps x (rebuilt slot, previous slot)
  float2* exch0 = stage0 + 8 * 2 * kStageSlot;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kInvBufs; ++i) {
      mbar_init(exch_full + i, kThreads / 32);
      mbar_init(exch_empty + i, kThreads / 32);
    }
    for (int i = 0; i < 8; ++i) mbar_init(stage_full + i, 1);
  }
  __syncthreads();
  const int first = blockIdx.x * P.tiles_per_cta;
  int n = P.n_tiles - first;
  if (n > P.tiles_per_cta) n = P.tiles_per_cta;
  if (n <= 0) return;
  TileCursor c = tile_cursor(first, P.tiles_per_clip);
  if (tid < kThreads) {
    const int warp = tid >> 5, lane = tid & 31;
    PairTw w;
    pair_tw_init(w, warp);
    float2* my_stage = stage0 + warp * 2 * kStageSlot;
    uint64_t* my_full = stage_full + warp;
    TileCursor c1 = c;
    if (lane == 0) inv_issue_stage_gl(P, maps.m, c1, warp, my_stage, my_full);
    tile_advance(c1, P.tiles_per_clip);
    int es = 0, use = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const int t0 = c.tt * P.g.FO - P.g.HL, t = t0 + lane;
      const bool live = (t >= 0 && t < P.n_frames);
      // lanes outside the clip read no |S| (their staged bins are TMA's zero fill): clamp the column, the result is dropped
      const int tc = live ? t : (t < 0 ? 0 : P.n_frames - 1);
      if (P.l2_prefetch && i + 1 < n) inv_prefetch_rows<4>(P.gl_mag, P, c1, warp, lane);       // |S| is the one array not staged
      mbar_wait(my_full, (uint32_t)(i & 1));
      InvLoadStagedGL load{my_stage + lane + (t0 & 1), nullptr, nullptr, P.gl_mag + (long long)c.b * kBins * P.T + tc, P.T,
                           P.gl_alpha, 0, 0, P, maps.m, c1, warp, my_stage, my_full, i + 1 < n};
      WaitBefore wb{exch_empty + es, (uint32_t)((use - 1) & 1), use >= 1};
      inv_stageA(exch0 + es * kExch, w, lane, warp, true, load, wb);
      mbar_arrive_warp(exch_full + es);
      tile_advance(c, P.tiles_per_clip);
      tile_advance(c1, P.tiles_per_clip);
      if (++es == P.n_bufs) { es = 0; ++use; }
    }
  } else {
    const int btid = tid - kThreads;
    LaneConst lc;
    lane_const_init(lc, tw_s, btid & 15);
    int es = 0, use = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      float2* exch = exch0 + es * kExch;
      mbar_wait(exch_full + es, (uint32_t)(use & 1));
      inv_phase1<kFast>(P, btid, c, exch, win_s, lc);
      named_bar_sync(1, kThreads);
      inv_phase2<kFast>(P, btid, c, exch, wtab);
      mbar_arrive_warp(exch_empty + es);
      tile_advance(c, P.tiles_per_clip);
      if (++es == P.n_bufs) { es = 0; ++use; }
    }
  }
}

struct GenericInvParams {
  InvParams P;
  int N, logN, F;
  float* frames;   // workspace [B, T, N]
};

__global__ void __launch_bounds__(256) istft_generic_frames_kernel(const GenericInvParams G) {
  extern __shared__ __align__(128) float smem[];
  float2* buf = reinterpret_cast<float2*>(smem);
  const InvParams& P = G.P;
  const int N = G.N;
  for (int fix = blockIdx.x; fix < P.n_tiles; fix += gridDim.x) {
    const int b = (int)(fix / P.n_frames);
    const int t = (int)(fix % P.n_frames);
    const bool db = P.db_flags ? (P.db_flags[b] != 0) : false;
    const long long base = (long long)b * G.F * P.T + t;
    for (int k = threadIdx.x; k < G.F; k += blockDim.x) {
      float xr, xi;
      inv_load_runtime(P, base + (long long)k * P.T, db, xr, xi);
      if (k == 0 || k == N / 2) xi = 0.0f;
      buf[__brev((unsigned)k) >> (32 - G.logN)] = make_float2(xr, xi);
      if (k != 0 && k != N / 2) buf[__brev((unsigned)(N - k)) >> (32 - G.logN)] = make_float2(xr, -xi);
    }
    __syncthreads();
    smem_fft(buf, N, G.logN, true);
    float* dst = G.frames + ((long long)b * P.T + t) * N;
    const float scale = 1.0f / (float)N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) dst[n] = buf[n].x * scale * P.window[n];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) istft_generic_ola_kernel(const GenericInvParams G) {
  const InvParams& P = G.P;
  const int N = G.N;
  for (int b = blockIdx.y; b < P.B; b += gridDim.y) {                     // clip per blockIdx.y: no division per sample
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < P.out_len; s += gridDim.x * blockDim.x) {
      const long long p = (long long)s + P.pad;
      long long f_lo = p - (N - 1) + P.hop - 1;
      f_lo = f_lo > 0 ? f_lo / P.hop : 0;
      long long f_hi = p / P.hop;
      if (f_hi > P.n_frames - 1) f_hi = P.n_frames - 1;
      float acc = 0.0f;
      for (long long f = f_lo; f <= f_hi; ++f)
        acc += G.frames[((long long)b * P.T + f) * N + (p - f * P.hop)];
      P.out[(long long)b * P.out_pitch + s] = acc * P.inv_wss[s];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------
// librosa.filters.window_sumsquare accumulates float32 x += float64 w^2 frame by frame; replayed here
// per sample in the same (increasing frame) order.
__global__ void inv_wss_kernel(const float* window, int N, int hop, int pad, int n_frames,
                               float* inv_wss, int out_len) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < out_len; s += gridDim.x * blockDim.x) {
    const long long p = (long long)s + pad;
    long long f_lo = p - (N - 1) + hop - 1;
    f_lo = f_lo > 0 ? f_lo / hop : 0;
    long long f_hi = p / hop;
    if (f_hi > n_frames - 1) f_hi = n_frames - 1;
    float acc = 0.0f;
    for (long long f = f_lo; f <= f_hi; ++f) {
      const double w = (double)window[p - f * hop];
      acc = (float)((double)acc + w * w);
    }
    inv_wss[s] = acc > kFltMin ? 1.0f / acc : 1.0f;
  }
}

// Griffin-Lim update (librosa.griffinlim loop body, utils.py:330-332), two bins per thread:
//   angles = rebuilt - alpha * tprev ; angles /= |angles| + tiny ; angles *= S ; tprev = rebuilt
// `spec` holds the rebuilt spectrum on entry and the new angles on exit.  Kept OUT of the forward kernel's
// epilogue on purpose: fused there, each bin's tprev / S loads sit on a dependent chain inside a 128-register
// thread (measured 20.6 ms per launch against 0.5 ms for the plain complex forward + 1.5 ms for this kernel).
__global__ void __launch_bounds__(256) gl_update_kernel(float4* __restrict__ spec, float4* __restrict__ tprev,
                                                        const float2* __restrict__ mag, long long n2, float alpha,
                                                        int has_prev) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 rb = spec[i];
    const float2 m = mag[i];
    float ax = rb.x, ay = rb.y, bx = rb.z, by = rb.w;
    if (has_prev) {
      const float4 tp = tprev[i];
      ax -= alpha * tp.x; ay -= alpha * tp.y; bx -= alpha * tp.z; by -= alpha * tp.w;
    }
    tprev[i] = rb;
    const float sa = 1.0f / (sqrtf(ax * ax + ay * ay) + kFltMin);      // librosa's order: normalise, then scale by |S|
    const float sb = 1.0f / (sqrtf(bx * bx + by * by) + kFltMin);
    spec[i] = make_float4((ax * sa) * m.x, (ay * sa) * m.x, (bx * sb) * m.y, (by * sb) * m.y);
  }
}

// Same update when the rebuilt spectra of consecutive iterations ping-pong between two buffers: tprev IS the previous
// iteration's output, so nothing is copied (7 instead of 9 array passes per iteration).
__global__ void __launch_bounds__(256) gl_update_pp_kernel(const float4* __restrict__ rebuilt, const float4* __restrict__ tprev,
                                                           const float2* __restrict__ mag, float4* __restrict__ angles,
                                                           long long n2, float alpha, int has_prev) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 rb = __ldcs(rebuilt + i);
    const float2 m = mag[i];
    float ax = rb.x, ay = rb.y, bx = rb.z, by = rb.w;
    if (has_prev) {
      const float4 tp = __ldcs(tprev + i);
      ax -= alpha * tp.x; ay -= alpha * tp.y; bx -= alpha * tp.z; by -= alpha * tp.w;
    }
    const float sa = 1.0f / (sqrtf(ax * ax + ay * ay) + kFltMin);
    const float sb = 1.0f / (sqrtf(bx * bx + by * by) + kFltMin);
    angles[i] = make_float4((ax * sa) * m.x, (ay * sa) * m.x, (bx * sb) * m.y, (by * sb) * m.y);
  }
}

__global__ void gl_update_pp_tail_kernel(const float2* rebuilt, const float2* tprev, const float* mag, float2* angles,
                                         long long i, float alpha, int has_prev) {
  const float2 rb = rebuilt[i];
  float ax = rb.x, ay = rb.y;
  if (has_prev) { const float2 tp = tprev[i]; ax -= alpha * tp.x; ay -= alpha * tp.y; }
  const float s = 1.0f / (sqrtf(ax * ax + ay * ay) + kFltMin);
  angles[i] = make_float2((ax * s) * mag[i], (ay * s) * mag[i]);
}

__global__ void gl_update_tail_kernel(float2* spec, float2* tprev, const float* mag, long long i, float alpha, int has_prev) {
  const float2 rb = spec[i];
  float ax = rb.x, ay = rb.y;
  if (has_prev) { const float2 tp = tprev[i]; ax -= alpha * tp.x; ay -= alpha * tp.y; }
  tprev[i] = rb;
  const float s = 1.0f / (sqrtf(ax * ax + ay * ay) + kFltMin);
  spec[i] = make_float2((ax * s) * mag[i], (ay * s) * mag[i]);
}

// librosa.griffinlim(init="random"): angles = exp(2 pi j U[0, 1)) (np.random.default_rng: any uniform stream will do; the reference
// never seeds it).  Philox4x32-10 (Salmon et al., SC'11) counter RNG: thread i turns counter (i, 0, 0, 0) under key `seed` into four
// 32-bit words = four phases; nothing is read, the only traffic is the 8 bytes per bin written.
__device__ __forceinline__ void philox_round(unsigned (&c)[4], unsigned k0, unsigned k1) {
  const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__global__ void __launch_bounds__(256) random_phasors_kernel(float2* angles, long long n, unsigned long long seed) {
  const long long quads = (n + 3) >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (long long)gridDim.x * blockDim.x) {
    unsigned c[4] = {(unsigned)i, (unsigned)(i >> 32), 0u, 0u};
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = 4 * i + j;
      if (e < n) {
        float sn, cs;
        sincospif((float)(c[j] >> 8) * (2.0f / 16777216.0f), &sn, &cs);      // phase 2 pi u, u = 24 random bits / 2^24
        angles[e] = make_float2(cs, sn);
      }
    }
  }
}

__global__ void scale_angles_kernel(float2* angles, const float* mag, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float m = mag[i];
    float2 a = angles[i];
    a.x *= m; a.y *= m;
    angles[i] = a;
  }
}

// librosa.griffinlim handed a COMPLEX "magnitude" S (tests/utils_test.py:624-645): every `angles *= S` is a complex product.
__global__ void scale_angles_c_kernel(float2* angles, const float2* spec, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 a = angles[i], s = spec[i];
    angles[i] = make_float2(a.x * s.x - a.y * s.y, a.x * s.y + a.y * s.x);
  }
}

__global__ void __launch_bounds__(256) gl_update_pp_c_kernel(const float2* __restrict__ rebuilt, const float2* __restrict__ tprev,
                                                             const float2* __restrict__ spec, float2* __restrict__ angles,
                                                             long long n, float alpha, int has_prev) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 rb = __ldcs(rebuilt + i);
    float ax = rb.x, ay = rb.y;
    if (has_prev) { const float2 tp = __ldcs(tprev + i); ax -= alpha * tp.x; ay -= alpha * tp.y; }
    const float inv = 1.0f / (sqrtf(ax * ax + ay * ay) + kFltMin);
    ax *= inv; ay *= inv;
    const float2 s = spec[i];
    angles[i] = make_float2(ax * s.x - ay * s.y, ax * s.y + ay * s.x);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// complex spectrogram [B, 257, T] as the float32 tensor (2t, p, j, b), bin = 16 j + p; box = 64 floats x 1 x rows_j x 1
static bool inv_make_map(CUtensorMap* map, const float2* spec, int B, int T, int n_frames, int rows_j) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[4] = {2ull * (cuuint64_t)n_frames, 16ull, 17ull, (cuuint64_t)B};
  const cuuint64_t strides[3] = {8ull * T, 128ull * T, 8ull * kBins * T};          // bytes, dims 1..3
  const cuuint32_t box[4] = {2u * kStageFr, 1u, (cuuint32_t)rows_j, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float2*>(spec), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool inv_tma_ok(const InvParams& P) {
  if (!P.spec || P.gl_mag || (P.T & 1) || (reinterpret_cast<uintptr_t>(P.spec) & 15)) return false;
  // On by default (AIP_INV_TMA=0 runs the direct-load kernel): 0.479 ms against 0.524 ms (1024 x 10 s, hop 192), 0.681 against
  // 0.730 ms at hop 128 -- one staging slot per warp next to the ring of two exchange buffers (profiles/README.md).
  return tunables().inv_tma != 0;
}

static bool inv_gl_tma_ok(const InvParams& P) {
  if (!P.spec || !P.gl_mag || !P.gl_prev || (P.T & 1)) return false;
  if ((reinterpret_cast<uintptr_t>(P.spec) & 15) || (reinterpret_cast<uintptr_t>(P.gl_prev) & 15)) return false;
  return tunables().inv_tma != 0;
}

static void inv_fill_ola(InvParams& P) {
  P.hop_magic = (unsigned)((0x100000000ULL + (unsigned)P.hop - 1) / (unsigned)P.hop);
  P.col_magic = (unsigned)((0x100000000ULL + (unsigned)(P.hop / 2) - 1) / (unsigned)(P.hop / 2));
  P.ola_terms = (kNfft + P.hop - 1) / P.hop;
  P.ola_dq = (2 * kThreads) / P.hop;
  P.ola_dr = (2 * kThreads) % P.hop;
  // reference frame whose K-1 predecessors exist and whose whole hop lies inside the output
  int f_ref = P.ola_terms - 1;
  const int need = (P.pad + P.hop - 1) / P.hop;
  if (f_ref < need) f_ref = need;
  const long long s_ref = (long long)f_ref * P.hop - P.pad;
  P.wss_ref = (f_ref <= P.n_frames - 1 && s_ref >= 0 && s_ref + P.hop <= P.out_len) ? (int)s_ref : -1;
}

static bool inv_fast_ok(const aip_stft_desc* d) {
  if (d->n_fft != 512 || (d->hop & 1)) return false;
  return inv_geom(d->hop, d->center ? 256 : 0).FO >= 4;
}

static int run_inv(const aip_stft_desc* desc, InvParams P, long long length, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!desc || !desc->window || !P.out || !P.inv_wss || (!P.spec && !P.mag)) return AIP_ERR_ARG;
  if (!is_pow2(desc->n_fft) || desc->n_fft < 32 || desc->n_fft > 4096 || desc->hop <= 0) return AIP_ERR_UNSUPPORTED;
  if (P.B < 0 || P.T < 1 || length < 0) return AIP_ERR_ARG;
  if (P.mag_domain < DOM_LINEAR || P.mag_domain > DOM_EXPM1) return AIP_ERR_ARG;
  const long long out_len = istft_length(P.T, desc->n_fft, desc->hop, desc->center, length);
  if (out_len < 0 || P.out_pitch < out_len) return AIP_ERR_ARG;
  if (P.B == 0 || out_len == 0) return AIP_OK;
  P.hop = desc->hop;
  P.pad = desc->center ? desc->n_fft / 2 : 0;
  P.n_frames = (int)istft_used_frames(P.T, desc->n_fft, desc->hop, desc->center, length);
  P.out_len = (int)out_len;
  P.window = desc->window;
  cudaError_t e;
  if (inv_fast_ok(desc)) {
    P.g = inv_geom(P.hop, P.pad);
    const long long span = (long long)P.g.FO * P.hop;
    P.tiles_per_clip = (int)((out_len + span - 1) / span);
    if ((long long)P.B * P.tiles_per_clip > 0x7fffffffLL || P.T > (1 << 22)) return AIP_ERR_UNSUPPORTED;
    P.n_tiles = (int)((long long)P.B * P.tiles_per_clip);
    P.vec_ok = ((P.out_pitch & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 7) == 0) &&
               ((reinterpret_cast<uintptr_t>(P.inv_wss) & 7) == 0);
    inv_fill_ola(P);
    P.l2_prefetch = tunables().inv_l2_prefetch;
    P.ola_fast = inv_ola_fast_kind(P.hop, P.pad, desc->win_length);
    if (P.ola_fast > 0 && !((tunables().ola_fast_mask >> (P.ola_fast - 1)) & 1)) P.ola_fast = 0;
    P.n_bufs = kInvBufsDefault;
    if (const int v = tunables().inv_bufs) { if (v >= 1 && v <= kInvBufs) P.n_bufs = v; }
    int grid = di.sms;
    if (grid > P.n_tiles) grid = P.n_tiles;
    P.tiles_per_cta = (P.n_tiles + grid - 1) / grid;
    grid = (P.n_tiles + P.tiles_per_cta - 1) / P.tiles_per_cta;
    if (inv_gl_tma_ok(P)) {
      GlMaps gm;
      if (inv_make_map(&gm.m[0], P.spec, P.B, P.T, P.n_frames, 16) && inv_make_map(&gm.m[1], P.spec, P.B, P.T, P.n_frames, 1) &&
          inv_make_map(&gm.m[2], P.gl_prev, P.B, P.T, P.n_frames, 16) && inv_make_map(&gm.m[3], P.gl_prev, P.B, P.T, P.n_frames, 1)) {
        P.n_bufs = 1;      // 8 x 2 staging slots (142 KB) + one exchange buffer (66 KB)
        const size_t smem_gl = (size_t)8 * 2 * kStageBytesWarp + (size_t)kExch * sizeof(float2);
        auto gk = P.ola_fast == 1 ? istft512_gl_tma_kernel<1> : (P.ola_fast == 2 ? istft512_gl_tma_kernel<2> : istft512_gl_tma_kernel<0>);
        e = cudaFuncSetAttribute(gk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gl);
        if (e != cudaSuccess) return (int)e;
        gk<<<(unsigned)grid, kFwdThreads, smem_gl, st>>>(P, gm);
        return (int)cudaGetLastError();
      }
    }
    if (inv_tma_ok(P)) {
      CUtensorMap map16, map1;
      if (inv_make_map(&map16, P.spec, P.B, P.T, P.n_frames, 16) && inv_make_map(&map1, P.spec, P.B, P.T, P.n_frames, 1)) {
        // 8 staging slots (71 KB) + the ring of exchange buffers (2 x 66 KB)
        if (P.n_bufs > 2) P.n_bufs = 2;
        const size_t smem_tma = (size_t)8 * kStageBytesWarp + (size_t)P.n_bufs * kExch * sizeof(float2);
        auto tk = P.ola_fast == 1 ? istft512_tma_kernel<1> : (P.ola_fast == 2 ? istft512_tma_kernel<2> : istft512_tma_kernel<0>);
        e = cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
        if (e != cudaSuccess) return (int)e;
        tk<<<(unsigned)grid, kFwdThreads, smem_tma, st>>>(P, map16, map1);
        return (int)cudaGetLastError();
      }
    }
    const size_t smem = (size_t)P.n_bufs * kExch * sizeof(float2);
    void (*kern)(const InvParams) = nullptr;
#define AIP_INV_CASE(M) \
    case (M): kern = P.ola_fast == 1 ? istft512_kernel<(M), 1> : (P.ola_fast == 2 ? istft512_kernel<(M), 2> : istft512_kernel<(M), 0>); break;
    switch (inv_mode_of(P)) {
      AIP_INV_CASE(INV_SPEC)
      AIP_INV_CASE(inv_mag_mode(0, false))
      AIP_INV_CASE(inv_mag_mode(0, true))
      AIP_INV_CASE(inv_mag_mode(1, false))
      AIP_INV_CASE(inv_mag_mode(1, true))
      AIP_INV_CASE(inv_mag_mode(2, false))
      AIP_INV_CASE(INV_BLEND)
      AIP_INV_CASE(INV_BLEND_LIN)
      AIP_INV_CASE(INV_BLEND_EXPM1)
      AIP_INV_CASE(INV_GL)
      default: kern = P.ola_fast == 1 ? istft512_kernel<inv_mag_mode(2, true), 1>
                                      : (P.ola_fast == 2 ? istft512_kernel<inv_mag_mode(2, true), 2> : istft512_kernel<inv_mag_mode(2, true), 0>);
               break;
    }
#undef AIP_INV_CASE
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<(unsigned)grid, kFwdThreads, smem, st>>>(P);
    e = cudaGetLastError();
  } else if (pow2_ola_ok(desc->n_fft, desc->hop)) {
    e = launch_inv_pow2(P, desc->n_fft, nullptr, di, st);
  } else {
    const size_t need = aip_istft_workspace_bytes(desc, P.B, P.T);
    if (!workspace || workspace_bytes < need) return AIP_ERR_WORKSPACE;
    GenericInvParams G;
    G.N = desc->n_fft; G.logN = ilog2(desc->n_fft); G.F = desc->n_fft / 2 + 1;
    G.frames = static_cast<float*>(workspace);
    if ((long long)P.B * P.n_frames > 0x7fffffffLL) return AIP_ERR_UNSUPPORTED;
    P.n_tiles = (int)((long long)P.B * P.n_frames);
    G.P = P;
    if (pow2_ok(desc->n_fft)) {
      e = launch_inv_pow2(P, desc->n_fft, G.frames, di, st);
    } else {
      const size_t smem = (size_t)desc->n_fft * sizeof(float2);
      e = cudaFuncSetAttribute(istft_generic_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      long long grid = (long long)di.sms * 8;
      if (grid > P.n_tiles) grid = P.n_tiles;
      istft_generic_frames_kernel<<<(unsigned)grid, 256, smem, st>>>(G);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) return (int)e;
    {
      long long gx = ((long long)P.out_len + 2047) / 2048;
      if (gx > 1024) gx = 1024;
      istft_generic_ola_kernel<<<dim3((unsigned)gx, (unsigned)(P.B < 65535 ? P.B : 65535)), 256, 0, st>>>(G);
    }
    e = cudaGetLastError();
  }
  return (int)e;
}

}  // namespace aip

using namespace aip;

extern "C" {

int aip_istft_blend_f32(const aip_stft_desc* desc, const float* model_out, const float* blend_in, const float* blend_mask,
                        const float* phase, int32_t mag_domain, int64_t B, int64_t T, int64_t length,
                        const float* inv_wss, float* wave_out, int64_t out_pitch, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (mag_domain != DOM_POW10 && mag_domain != DOM_DB) return AIP_ERR_UNSUPPORTED;
  return aip_istft_handoff_f32(desc, model_out, blend_in, blend_mask, 0, phase, mag_domain, B, T, length, inv_wss, wave_out,
                               out_pitch, nullptr, nullptr, 0, workspace, workspace_bytes, stream);
}

// inverse with the per-clip peak taken in its overlap-add, then ONE pass over the waveform: the in-place scaling of
// librosa.util.normalize, or -- pcm_out given -- the same division followed by the float -> PCM_16 conversion, written to pcm_out
// (P.out then keeps the un-normalised waveform)
static int normalized_tail(const aip_stft_desc* desc, InvParams& P, int64_t length, long long out_len, float* peaks,
                           int16_t* pcm_out, int64_t pcm_pitch, void* workspace, size_t workspace_bytes, const DevInfo& di,
                           cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(peaks, 0, (size_t)P.B * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const bool fused = inv_fast_ok(desc);
  P.peaks = fused ? peaks : nullptr;
  const int rc = run_inv(desc, P, length, workspace, workspace_bytes, st);
  if (rc != AIP_OK) return rc;
  if (out_len <= 0) return AIP_OK;
  if (!fused) {      // generic n_fft path: separate peak pass
    if (P.B > 65535) return AIP_ERR_UNSUPPORTED;
    e = launch_peak(P.out, P.out_pitch, P.B, out_len, peaks, st);
    if (e != cudaSuccess) return (int)e;
  }
  if (pcm_out)
    return (int)launch_pcm16(P.out, P.out_pitch, reinterpret_cast<short*>(pcm_out), pcm_pitch, P.B, out_len, peaks, di.sms, st);
  return (int)launch_peak_scale(P.out, P.out_pitch, P.out, P.out_pitch, P.B, out_len, peaks, di.sms, st);
}

int aip_istft_handoff_f32(const aip_stft_desc* desc, const float* model_out, const float* blend_in, const float* blend_mask,
                          int32_t mask_keeps_input, const float* phase, int32_t mag_domain, int64_t B, int64_t T,
                          int64_t length, const float* inv_wss, float* wave_out, int64_t out_pitch, float* peaks,
                          int16_t* pcm_out, int64_t pcm_pitch, void* workspace, size_t workspace_bytes, void* stream) {
  if (B > 0x7fffffffLL || T > 0x7fffffffLL || B < 0) return AIP_ERR_ARG;
  if (!desc || !model_out || !blend_in || !blend_mask || !phase) return AIP_ERR_ARG;
  if (mag_domain < DOM_LINEAR || mag_domain > DOM_EXPM1) return AIP_ERR_ARG;
  if (B == 0) return AIP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  InvParams P{};
  // m = mag * g + blend_in * (1 - g) in the loader: the GAN convention (mask 1 OUTSIDE the gap) is the same expression with
  // the two magnitude arrays exchanged
  P.mag = mask_keeps_input ? blend_in : model_out;
  P.blend_in = mask_keeps_input ? model_out : blend_in;
  P.blend_mask = blend_mask; P.phase = phase; P.mag_domain = mag_domain;
  P.B = (int)B; P.T = (int)T; P.inv_wss = inv_wss; P.out = wave_out; P.out_pitch = out_pitch;
  const long long out_len = istft_length(T, desc->n_fft, desc->hop, desc->center, length);
  if (pcm_out && pcm_pitch < out_len) return AIP_ERR_ARG;
  if (!peaks) {
    const int rc = run_inv(desc, P, length, workspace, workspace_bytes, st);
    if (rc != AIP_OK || !pcm_out || out_len <= 0) return rc;
    return (int)launch_pcm16(wave_out, out_pitch, reinterpret_cast<short*>(pcm_out), pcm_pitch, B, out_len, nullptr, di.sms, st);
  }
  return normalized_tail(desc, P, length, out_len, peaks, pcm_out, pcm_pitch, workspace, workspace_bytes, di, st);
}

size_t aip_istft_workspace_bytes(const aip_stft_desc* desc, int64_t B, int64_t T) {
  if (!desc || B <= 0 || T <= 0 || desc->n_fft <= 0 || inv_fast_ok(desc) || pow2_ola_ok(desc->n_fft, desc->hop)) return 0;
  return (size_t)B * (size_t)T * (size_t)desc->n_fft * sizeof(float);
}

int aip_istft_f32(const aip_stft_desc* desc, const float* spec, const float* mag, const float* phase,
                  int32_t mag_domain, const int32_t* db_flags, int64_t B, int64_t T, int64_t length,
                  const float* inv_wss, float* wave_out, int64_t out_pitch, void* workspace,
                  size_t workspace_bytes, void* stream) {
  if (B > 0x7fffffffLL || T > 0x7fffffffLL) return AIP_ERR_ARG;
  InvParams P{};
  P.spec = reinterpret_cast<const float2*>(spec); P.mag = mag; P.phase = phase; P.mag_domain = mag_domain;
  P.db_flags = db_flags; P.B = (int)B; P.T = (int)T; P.inv_wss = inv_wss; P.out = wave_out; P.out_pitch = out_pitch;
  return run_inv(desc, P, length, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int aip_istft_normalized_f32(const aip_stft_desc* desc, const float* spec, const float* mag, const float* phase,
                             int32_t mag_domain, const int32_t* db_flags, int64_t B, int64_t T, int64_t length,
                             const float* inv_wss, float* wave_out, int64_t out_pitch, float* peaks, int16_t* pcm_out,
                             int64_t pcm_pitch, void* workspace, size_t workspace_bytes, void* stream) {
  if (B > 0x7fffffffLL || T > 0x7fffffffLL || B < 0 || !peaks || !desc) return AIP_ERR_ARG;
  if (B == 0) return AIP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  const long long out_len = istft_length(T, desc->n_fft, desc->hop, desc->center, length);
  if (pcm_out && pcm_pitch < out_len) return AIP_ERR_ARG;
  InvParams P{};
  P.spec = reinterpret_cast<const float2*>(spec); P.mag = mag; P.phase = phase; P.mag_domain = mag_domain;
  P.db_flags = db_flags; P.B = (int)B; P.T = (int)T; P.inv_wss = inv_wss; P.out = wave_out; P.out_pitch = out_pitch;
  return normalized_tail(desc, P, length, out_len, peaks, pcm_out, pcm_pitch, workspace, workspace_bytes, di, st);
}

int aip_inv_window_sumsquare_f32(const aip_stft_desc* desc, int64_t T, int64_t length, float* inv_wss,
                                 int64_t out_len, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!desc || !desc->window || !inv_wss || T < 1 || desc->hop <= 0 || desc->n_fft <= 0) return AIP_ERR_ARG;
  if (out_len != istft_length(T, desc->n_fft, desc->hop, desc->center, length)) return AIP_ERR_ARG;
  if (out_len == 0) return AIP_OK;
  const int nf = (int)istft_used_frames(T, desc->n_fft, desc->hop, desc->center, length);
  inv_wss_kernel<<<ew_grid(out_len, di.sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      desc->window, desc->n_fft, desc->hop, desc->center ? desc->n_fft / 2 : 0, nf, inv_wss, (int)out_len);
  return (int)cudaGetLastError();
}

int aip_random_phasors_f32(float* angles, int64_t n, uint64_t seed, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!angles || n < 0 || (reinterpret_cast<uintptr_t>(angles) & 7)) return AIP_ERR_ARG;
  if (n == 0) return AIP_OK;
  random_phasors_kernel<<<ew_grid((n + 3) / 4, di.sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float2*>(angles), n,
                                                                                                   seed);
  return (int)cudaGetLastError();
}

int aip_griffinlim_f32(const aip_stft_desc* desc, const float* mag, float* angles, float* tprev, int64_t B,
                       int64_t T, int32_t n_iter, float momentum, const float* inv_wss, float* wave_out,
                       int64_t out_pitch, void* workspace, size_t workspace_bytes, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!desc || !mag || !angles || !tprev || !wave_out || n_iter < 0 || momentum < 0.0f) return AIP_ERR_ARG;
  if (B > 0x7fffffffLL || T > 0x7fffffffLL || B < 0 || T < 1) return AIP_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(angles) & 15) || (reinterpret_cast<uintptr_t>(tprev) & 15) ||
      (reinterpret_cast<uintptr_t>(mag) & 7)) return AIP_ERR_ARG;     // vector access of the state arrays
  if (B == 0) return AIP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long F = desc->n_fft / 2 + 1;
  const long long n = (long long)B * F * T;
  const long long out_len = istft_length(T, desc->n_fft, desc->hop, desc->center, 0);
  if (num_frames(out_len, desc->n_fft, desc->hop, desc->center) != T) return AIP_ERR_UNSUPPORTED;
  scale_angles_kernel<<<ew_grid(n, di.sms), 256, 0, st>>>(reinterpret_cast<float2*>(angles), mag, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  InvParams I{};
  I.spec = reinterpret_cast<const float2*>(angles); I.B = (int)B; I.T = (int)T; I.inv_wss = inv_wss;
  I.out = wave_out; I.out_pitch = out_pitch;
  // With room for one more [B,F,T] complex array behind the istft workspace the rebuilt spectra ping-pong between
  // `tprev` and that array and the update reads the previous one in place of a copy.
  const size_t ws_istft = (aip_istft_workspace_bytes(desc, B, T) + 15) & ~(size_t)15;
  const size_t ws_pp = (size_t)n * sizeof(float2);
  char* ws_base = static_cast<char*>(workspace);
  float2* rb[2] = {reinterpret_cast<float2*>(tprev), nullptr};
  if (workspace && workspace_bytes >= ws_istft + ws_pp && ((reinterpret_cast<uintptr_t>(ws_base + ws_istft) & 15) == 0))
    rb[1] = reinterpret_cast<float2*>(ws_base + ws_istft);
  const bool pingpong = rb[1] != nullptr;
  const float alpha = momentum / (1.0f + momentum);
  const long long n2 = n / 2;
  long long g = (n2 + 255) / 256;
  if (g > (long long)di.sms * 32) g = (long long)di.sms * 32;
  // With the two ping-pong buffers the phase update needs no kernel and no `angles` array of its own: the inverse kernel of
  // the NEXT iteration (or the final one) reads rebuilt[it], rebuilt[it - 1] and |S| and projects while it loads (INV_GL,
  // InvLoadGL) -- 3 array passes fewer per iteration.  The forward kernel of iteration `it` overwrites rebuilt[it - 2], which
  // the inverse kernel before it on the stream was the last to read.  Needs the fast n_fft = 512 path.
  // ... or the tiled power-of-two kernels, whose transposed load does the same (istft_pow2_kernel).
  const bool fused = pingpong && ((fwd_fast_ok(desc, di) && inv_fast_ok(desc)) || pow2_ok(desc->n_fft)) && !tunables().gl_unfused;
  for (int it = 0; it < n_iter; ++it) {
    int rc = run_inv(desc, I, 0, workspace, workspace_bytes, st);
    if (rc != AIP_OK) return rc;
    float2* rebuilt = pingpong ? rb[it & 1] : reinterpret_cast<float2*>(angles);
    FwdParams P{};
    P.wave = wave_out; P.wave_pitch = out_pitch; P.B = (int)B; P.L = (int)out_len;
    P.mag_kind = MAG_NONE;
    P.spec = rebuilt;
    rc = run_fwd(desc, P, T, st);
    if (rc != AIP_OK) return rc;
    if (fused) {
      I.spec = rebuilt; I.gl_mag = mag;
      I.gl_prev = it > 0 ? rb[(it + 1) & 1] : rebuilt; I.gl_alpha = it > 0 ? alpha : 0.0f;      // librosa: tprev is None at first
    } else if (pingpong) {
      const float2* prev = rb[(it + 1) & 1];
      if (n2 > 0)
        gl_update_pp_kernel<<<(unsigned)g, 256, 0, st>>>(reinterpret_cast<const float4*>(rebuilt), reinterpret_cast<const float4*>(prev),
                                                         reinterpret_cast<const float2*>(mag), reinterpret_cast<float4*>(angles), n2,
                                                         alpha, it > 0);
      if (n & 1)
        gl_update_pp_tail_kernel<<<1, 1, 0, st>>>(rebuilt, prev, mag, reinterpret_cast<float2*>(angles), n - 1, alpha, it > 0);
    } else {
      if (n2 > 0)
        gl_update_kernel<<<(unsigned)g, 256, 0, st>>>(reinterpret_cast<float4*>(angles), reinterpret_cast<float4*>(tprev),
                                                      reinterpret_cast<const float2*>(mag), n2, alpha, it > 0);
      if (n & 1)
        gl_update_tail_kernel<<<1, 1, 0, st>>>(reinterpret_cast<float2*>(angles), reinterpret_cast<float2*>(tprev), mag,
                                               n - 1, alpha, it > 0);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return run_inv(desc, I, 0, workspace, workspace_bytes, st);
}

int aip_griffinlim_c64_f32(const aip_stft_desc* desc, const float* spec, float* angles, float* tprev, int64_t B,
                           int64_t T, int32_t n_iter, float momentum, const float* inv_wss, float* wave_out,
                           int64_t out_pitch, void* workspace, size_t workspace_bytes, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!desc || !spec || !angles || !tprev || !wave_out || n_iter < 0 || momentum < 0.0f) return AIP_ERR_ARG;
  if (B > 0x7fffffffLL || T > 0x7fffffffLL || B < 0 || T < 1) return AIP_ERR_ARG;
  if (B == 0) return AIP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long F = desc->n_fft / 2 + 1;
  const long long n = (long long)B * F * T;
  const long long out_len = istft_length(T, desc->n_fft, desc->hop, desc->center, 0);
  if (num_frames(out_len, desc->n_fft, desc->hop, desc->center) != T) return AIP_ERR_UNSUPPORTED;
  const size_t ws_istft = (aip_istft_workspace_bytes(desc, B, T) + 15) & ~(size_t)15;
  if (!workspace || workspace_bytes < ws_istft + (size_t)n * sizeof(float2)) return AIP_ERR_WORKSPACE;
  float2* rb[2] = {reinterpret_cast<float2*>(tprev), reinterpret_cast<float2*>(static_cast<char*>(workspace) + ws_istft)};
  const float2* S = reinterpret_cast<const float2*>(spec);
  float2* ang = reinterpret_cast<float2*>(angles);
  const int grid = ew_grid(n, di.sms);
  scale_angles_c_kernel<<<grid, 256, 0, st>>>(ang, S, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  InvParams I{};
  I.spec = ang; I.B = (int)B; I.T = (int)T; I.inv_wss = inv_wss; I.out = wave_out; I.out_pitch = out_pitch;
  const float alpha = momentum / (1.0f + momentum);
  for (int it = 0; it < n_iter; ++it) {
    int rc = run_inv(desc, I, 0, workspace, workspace_bytes, st);
    if (rc != AIP_OK) return rc;
    FwdParams P{};
    P.wave = wave_out; P.wave_pitch = out_pitch; P.B = (int)B; P.L = (int)out_len;
    P.mag_kind = MAG_NONE;
    P.spec = rb[it & 1];
    rc = run_fwd(desc, P, T, st);
    if (rc != AIP_OK) return rc;
    gl_update_pp_c_kernel<<<grid, 256, 0, st>>>(rb[it & 1], rb[(it + 1) & 1], S, ang, n, alpha, it > 0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return run_inv(desc, I, 0, workspace, workspace_bytes, st);
}

}  // extern "C"
