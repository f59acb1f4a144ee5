// sm_100a kernels + C ABI (include/aip_b200.h) of the spectrogram front-end / back-end.
//
// Kernels
//   stft512_fwd_kernel   persistent; tile = 32 frames of one clip; waveform staged once in shared
//                        memory by TMA bulk copies, register 16x16 FFT (zero window taps pruned), packed
//                        split pass, fused |S| / log / phase / mask epilogues, coalesced [F,T] stores.
//   istft512_kernel      persistent; tile = 32 frames -> FO hops of output; split-pass prologue,
//                        register inverse FFT, synthesis window, overlap-add and window-sum-square
//                        normalisation out of shared memory (no global atomics, halo frames recomputed),
//                        optional per-clip peak for the fused normalisation.
//   istft512_tma_kernel  the same with stage A's rows staged by 4-D TMA tensor boxes (switch AIP_INV_TMA=1).
//   gl_update*_kernel    the Griffin-Lim phase update between the two transforms.
//   gap variants         (aip_stft_gap_variants_f32: G gapped spectrograms per file from ONE clean transform)
//                        variant_meta_kernel -> variant_fill_tma_kernel (clean block staged in shared memory at the four
//                        16-byte phases, one bulk copy shared -> global per chunk and variant) -> stft512_fwd_kernel
//                        <mode | FWD_VARIANT> (re-transform of the one or two tiles a gap touches).
//   stft_generic_* / istft_generic_*   any power-of-two n_fft in [32, 4096] (or odd hop): one frame
//                        per CTA, shared-memory radix-2.  Correct, not tuned: the reference's
//                        models only ever use n_fft = 512 (config.py:28, GAN/config.yaml:12).
//   small elementwise / reduction kernels for the gap, mask, dB-heuristic and peak-normalise rows.
#include <cuda.h>            // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>

#include "../../include/aip_b200.h"
#include "aip_tiles.cuh"

namespace aip {

// ---------------------------------------------------------------------------------------------------
// n_fft = 512 kernels
// ---------------------------------------------------------------------------------------------------
// ---- TMA (1-D bulk copy) + mbarrier plumbing --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// blocks until the phase with the given parity has completed (try_wait parks the warp for a hardware-
// defined time per attempt; a longer suspend-time hint measured SLOWER: later wake-ups)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if defined(AIP_MBAR_HINT_NS)
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(AIP_MBAR_HINT_NS) : "memory");
#elif defined(AIP_MBAR_SLEEP_NS)
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!ok) __nanosleep(AIP_MBAR_SLEEP_NS);
  } while (!ok);
#else
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}

// global -> shared bulk copy, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// global -> shared TMA tensor tile (SASS: UTMALDG), 4-D coordinates, completion on the mbarrier; elements outside the
// tensor are zero-filled and still counted in the transaction bytes
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

// shared -> global bulk copy (SASS: UBLKCP), tracked by the issuing thread's bulk async-group; 16-byte aligned on both sides
__device__ __forceinline__ void tma_store_1d(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one arrival per WARP (barrier counts are in warps): every arrival wakes the waiters, so 8 instead of 256
// arrivals per phase keeps the parked warps parked.  __syncwarp orders the other lanes' shared-memory
// accesses before lane 0's release.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one elected thread: start the bulk copy of a tile's in-range samples (or just complete the phase)
__device__ __forceinline__ void fwd_issue_tile(const FwdTilePlan& q, float* tile, uint64_t* bar) {
#if defined(AIP_ABLATE_LOADS)      // timing experiment only: no waveform traffic, stage 1 runs on whatever the buffer holds
  if (false) {
#else
  if (q.n_bulk > 0) {
#endif
    const uint32_t bytes = (uint32_t)q.n_bulk * 4u;
    mbar_expect_tx(bar, bytes);
    tma_load_1d(tile + q.v_lo, q.src + q.g0 + q.v_lo, bytes, bar);
  } else {
    mbar_arrive(bar);
  }
}

struct ArriveRelease {
  uint64_t* bar;
  __device__ __forceinline__ void operator()() const { mbar_arrive_warp(bar); }
};

constexpr int kFwdThreads = 2 * kThreads;   // 8 stage-2 (consumer) warps + 8 stage-1 (producer) warps
constexpr int kFwdTileBufs = 3;             // deepest ring of staged-waveform buffers

// Which warps play which role.  A warp runs on scheduler (warp & 3); bit q of the map says whether the warp in slot
// q = warp >> 2 of every scheduler is a consumer (stage 2 forward / stage A inverse), so each scheduler always hosts
// two warps of each role.  0b0011 = warps 0..7 consume, 8..15 produce.  (Five layouts measured 1.642-1.646 ms: the
// mapping does not matter; the switch stays for experiments.)
#ifndef AIP_ROLE_MAP
#define AIP_ROLE_MAP 0x3
#endif
struct WarpRole {
  bool consumer;
  int rtid;      // thread index 0..255 inside the role
};
__device__ __forceinline__ WarpRole warp_role(int tid) {
  const int w = tid >> 5, q = w >> 2, s = w & 3;
  const bool cons = (AIP_ROLE_MAP >> q) & 1;
  const int below = cons ? __popc(AIP_ROLE_MAP & ((1 << q) - 1)) : __popc(~AIP_ROLE_MAP & ((1 << q) - 1));
  WarpRole r;
  r.consumer = cons;
  r.rtid = ((below * 4 + s) << 5) | (tid & 31);
  return r;
}

// A thread's position in the CTA's dynamic tile schedule: chunk index, tiles left in the chunk, cursor of the next tile.
constexpr int kSchedRing = 8;     // chunks in flight between the publisher and the slowest reader (they are < 4 tiles apart)
struct TileFeed {
  int k, left;
  TileCursor c;
};
// picks up chunk f.k once it has been published; false = the batch is exhausted
__device__ __forceinline__ bool feed_next(const FwdParams& P, TileFeed& f, const int* sched_start, uint64_t* sched_bar) {
  const int slot = f.k & (kSchedRing - 1);
  mbar_wait(sched_bar + slot, (uint32_t)((f.k / kSchedRing) & 1));
  const int s = sched_start[slot];
  int cnt = P.n_tiles - s;
  if (cnt > P.chunk) cnt = P.chunk;
  if (cnt <= 0) return false;
  f.c = tile_cursor(s, P.tiles_per_clip);
  f.left = cnt;
  ++f.k;
  return true;
}

// Warp-specialised, persistent, one CTA per SM, tiles handed out dynamically in small contiguous chunks.  Tiles flow through
//   TMA bulk copy -> tile[slot] (ring of up to 3) -> stage-1 warps (lane = n1; window, 16-pt DFT, twiddle) -> exch[es]
//   -> stage-2 warps (lane = frame; 2 x 16-pt DFT, packed split pass, |.|/log epilogue) -> HBM
// with mbarrier hand-offs (tile_full / tile_empty / exch_full / exch_empty), so the copies of tiles i+1 and i+2,
// stage 1 of tile i+1 and stage 2 of tile i overlap, and each role keeps ITS constants in registers
// (stage 1: 16 W256 twiddles per lane, read once from a shared-memory table; stage 2: 16 W512 twiddles per warp).
template <int kMode, int kZP, int kT = 0>
__global__ void __launch_bounds__(kFwdThreads, 1) stft512_fwd_kernel(const FwdParams P) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2 * kFwdTileBufs + 4];
  __shared__ __align__(8) uint64_t sched_bar[kSchedRing];      // count 1: chunk k published
  __shared__ int sched_start[kSchedRing];                      // first tile of chunk k (>= n_tiles: the batch is done)
  __shared__ __align__(16) float win_s[kWinTable];
  __shared__ __align__(8) float2 tw_s[kTwTable];
  window_table_fill(win_s, P.window, 0.5f, threadIdx.x, blockDim.x);
  twiddle_table_fill(tw_s, threadIdx.x, blockDim.x);
  uint64_t* tile_full = bars;                          // [3] count 1 (+ tx bytes)
  uint64_t* tile_empty = bars + kFwdTileBufs;          // [3] count 8 (stage-1 warps)
  uint64_t* exch_full = bars + 2 * kFwdTileBufs;       // [2] count 8 (stage-1 warps)
  uint64_t* exch_empty = bars + 2 * kFwdTileBufs + 2;  // [2] count 8 (stage-2 warps)
  const int ntb = P.n_tile_bufs;
  float2* exch0 = reinterpret_cast<float2*>(smem + ntb * P.tile_floats);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kFwdTileBufs; ++i) {
      mbar_init(tile_full + i, 1);
      mbar_init(tile_empty + i, kThreads / 32);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(exch_full + i, kThreads / 32);
      mbar_init(exch_empty + i, kThreads / 32);
    }
    for (int i = 0; i < kSchedRing; ++i) mbar_init(sched_bar + i, 1);
  }
  __syncthreads();
  // Tiles are handed out dynamically, P.chunk contiguous tiles per atomicAdd on P.tile_counter: the SMs do not all run
  // at the same speed (static equal runs left the average SM idle for 3.4 % of the kernel).  The elected stage-1 thread
  // draws a chunk when its request cursor runs dry and publishes its first tile through sched_start[] / sched_bar[];
  // every other thread picks the chunks up in the same order when its own cursor runs dry.
  const WarpRole role = warp_role(tid);
  if (!role.consumer) {
    // ------------------------------------------------------------------ producers: stage 1
    // The staged waveform sits in a ring of ntb buffers; ntb - 1 bulk copies are in flight while a tile is being
    // transformed.  (With two buffers the stage-1 warps spent 11 % of their time waiting for the ONE copy in flight:
    // under this kernel's write-heavy traffic a 26 KB read takes about as long as a tile, see profiles/README.md.)
    const int ptid = role.rtid;
    LaneConst lc;
    lane_const_init(lc, tw_s, ptid & 15);
    TileFeed fn{0, 0, TileCursor{0, 0}};     // request cursor (thread ptid == 0 only): draws the chunks
    bool more = true;                        // the request cursor has not hit the end of the batch yet
    auto request = [&](float* buf, uint64_t* bar) {      // ptid == 0: start the copy of the next tile, if there is one
      if (fn.left == 0) {
        const int s = (int)atomicAdd(P.tile_counter, (unsigned)P.chunk);
        sched_start[fn.k & (kSchedRing - 1)] = s;
        mbar_arrive(sched_bar + (fn.k & (kSchedRing - 1)));          // release: publishes the chunk to the CTA
        int cnt = P.n_tiles - s;
        if (cnt > P.chunk) cnt = P.chunk;
        ++fn.k;
        if (cnt <= 0) { more = false; return; }
        fn.c = tile_cursor(s, P.tiles_per_clip);
        fn.left = cnt;
      }
      if (kMode & FWD_VARIANT) {
        const int2 m = *reinterpret_cast<const int2*>(P.var_meta + 4 * fn.c.b + 2);      // {frame base, wave row}
        fwd_issue_tile(fwd_tile_plan_var(P, fn.c, 0, 0, m.x, m.y), buf, bar);
      } else {
        fwd_issue_tile(fwd_tile_plan_gap(P, fn.c, 0, 0), buf, bar);
      }
      tile_advance(fn.c, P.tiles_per_clip);
      --fn.left;
    };
    if (ptid == 0) {
      for (int k = 0; k < (ntb > 1 ? ntb - 1 : 1) && more; ++k) request(smem + k * P.tile_floats, tile_full + k);
    }
    TileFeed f{0, 0, TileCursor{0, 0}};
    int slot = 0, use = 0;               // ring slot of tile i and how often it has been used before
    int gs = 0, ge = 0, gap_clip = -1;   // gap range of the clip the cursor is in
    int fb = 0, vrow = 0;                // FWD_VARIANT: first recomputed frame and wave row of that variant
#pragma unroll 1
    for (int i = 0;; ++i) {
      if (f.left == 0 && !feed_next(P, f, sched_start, sched_bar)) break;
      const TileCursor c = f.c;
      float* tile = smem + slot * P.tile_floats;
      if (kMode & FWD_VARIANT) {
        if (c.b != gap_clip) {      // one 16-byte load per variant: {gap start, gap end, frame base, wave row}
          const int4 m = *reinterpret_cast<const int4*>(P.var_meta + 4 * c.b);
          gs = m.x; ge = m.y; fb = m.z; vrow = m.w;
          gap_clip = c.b;
        }
      } else if (P.gap_samples && c.b != gap_clip) {
        gs = P.gap_samples[2 * c.b];
        ge = P.gap_samples[2 * c.b + 1];
        gap_clip = c.b;
      }
      const FwdTilePlan q = (kMode & FWD_VARIANT) ? fwd_tile_plan_var(P, c, gs, ge, fb, vrow) : fwd_tile_plan_gap(P, c, gs, ge);
      if (ntb > 1 && ptid == 0 && more) {
        // the next request goes into the slot that tile i - 1 has just left
        const int ns = slot == 0 ? ntb - 1 : slot - 1;
        if (i >= 1) mbar_wait(tile_empty + ns, (uint32_t)((slot == 0 ? use - 1 : use) & 1));
        request(smem + ns * P.tile_floats, tile_full + ns);
      }
      mbar_wait(tile_full + slot, (uint32_t)(use & 1));
      if ((kMode & FWD_VARIANT) && !fwd_needs_edge_fixup(q)) {
        fwd_gap_zero_own(q, P.hop, ptid, tile);        // every variant tile holds a gap: no CTA-wide barrier for it
        __syncwarp();
      } else if (fwd_needs_fixup(q)) {
        fwd_fixup(q, ptid, tile);
        named_bar_sync(1, kThreads);
      }
      const int es = i & 1;
      if (i >= 2) mbar_wait(exch_empty + es, (uint32_t)(((i >> 1) - 1) & 1));
      fwd_phase1<kZP>(P, ptid, tile, exch0 + es * kExch, win_s, lc);
      mbar_arrive_warp(exch_full + es);
      fence_proxy_async();
      mbar_arrive_warp(tile_empty + slot);
      if (ntb == 1 && ptid == 0 && more) {
        mbar_wait(tile_empty, (uint32_t)(i & 1));
        request(smem, tile_full);
      }
      tile_advance(f.c, P.tiles_per_clip);
      --f.left;
      if (++slot == ntb) { slot = 0; ++use; }
    }
  } else {
    // ------------------------------------------------------------------ consumers: stage 2 + epilogue
    const int ctid = role.rtid;
    PairTw w;
    pair_tw_init(w, ctid >> 5);
    TileFeed f{0, 0, TileCursor{0, 0}};
    int var_clip = -1, fb_cur = 0;     // FWD_VARIANT: frame base of the variant the cursor is in
#pragma unroll 1
    for (int i = 0;; ++i) {
      if (f.left == 0 && !feed_next(P, f, sched_start, sched_bar)) break;
      const int es = i & 1;
      int fb = 0;
      if (kMode & FWD_VARIANT) {
        if (f.c.b != var_clip) { fb_cur = P.var_meta[4 * f.c.b + 2]; var_clip = f.c.b; }
        fb = fb_cur;
        // A variant tile writes 257 isolated, 4-byte-phased 128-byte row segments into a spectrogram that the copy pass wrote
        // a while ago: both ends of every segment are partial 32-byte sectors of lines that have left L2, and L2 fetches them
        // before it can merge the store -- with the stores 221 us, without them 117 us for 6400 tiles (ablation build).  So the
        // two lines of every row segment are requested from L2 NOW, while this warp would wait for stage 1 anyway (thread r
        // takes row r): 0.210 -> 0.176 ms.  Requesting them a whole tile ahead measured the same (0.173-0.175 ms).
        if (P.var_prefetch) {
          const int t0 = fb + f.c.tt * kFR;
          const int t1 = (t0 + kFR - 1) < P.T_out ? (t0 + kFR - 1) : (P.T_out - 1);
          const float* row = P.mag + ((long long)f.c.b * kBins + ctid) * P.T_out;
          prefetch_l2(row + t0);
          prefetch_l2(row + t1);
          if (ctid == 0) { prefetch_l2(row + 256LL * P.T_out + t0); prefetch_l2(row + 256LL * P.T_out + t1); }
        }
      }
      mbar_wait(exch_full + es, (uint32_t)((i >> 1) & 1));
      ArriveRelease rel{exch_empty + es};
      fwd_phase2<kMode, ArriveRelease, kT>(P, ctid, f.c, exch0 + es * kExch, w, rel, fb);
      tile_advance(f.c, P.tiles_per_clip);
      --f.left;
    }
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
// packed codelet leave no room for more), which bounds it by HBM latency.  Here every stage-A warp owns a private,
// double-buffered staging area that TMA fills one tile ahead: the spectrogram is described to the TMA unit as the
// 4-D tensor (2t, p, j, b) with bin k = 16 j + p, so the 16 rows p + 16 j a pair-job reads are ONE box of
// 68 floats x 1 x 16 x 1, the 16 partner rows 256 - p - 16 j a second one, and frames outside [0, n_frames) arrive
// as zeros.  Needs an even T (row pitch 8 T bytes must be a multiple of 16) and a 16-byte aligned base.
// The TMA unit needs the innermost start coordinate 16-byte aligned (an odd first frame raises an illegal-instruction
// fault, tools/microbench/tma3d_probe.cu), so a box starts at the even frame below t0 and is 34 frames wide.
constexpr int kStageFr = kFR + 2;                       // frames per staged row
constexpr int kStageB = 592;                            // float2 offset of region B (128-byte aligned: 4736 B)
constexpr int kStageSlot = 1136;                        // float2 per slot (9088 B): A = rows 0..16 (17 x 34), B = 16 rows
constexpr int kStageBytesWarp = 2 * kStageSlot * 8;     // two slots

struct InvLoadStaged {      // stage-A loader out of the warp's staging slot (lane = frame)
  const float2* slot;       // + lane + (t0 & 1)
  const float2* plo;
  const float2* phi;
  __device__ __forceinline__ void rows(int k_lo, int k_hi) {
    if (k_hi == 256) { plo = slot; phi = slot + 16 * kStageFr; }                               // job 0: rows j and 16 - j of region A
    else if (k_lo == 8) { plo = slot + kStageB; phi = slot + kStageB + 15 * kStageFr; }       // job 8: region B
    else { plo = slot; phi = slot + kStageB + 15 * kStageFr; }                                 // p > 0: A = p + 16 j, B = (16-p) + 16 j'
  }
  __device__ __forceinline__ void lo(int j, float& xr, float& xi) const { const float2 v = plo[j * kStageFr]; xr = v.x; xi = v.y; }
  __device__ __forceinline__ void hi(int j, float& xr, float& xi) const { const float2 v = phi[-j * kStageFr]; xr = v.x; xi = v.y; }
};

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

struct StagedBefore {       // runs once the slot has been read into registers: refill it, then wait for the exchange buffer
  const InvParams& P;
  const CUtensorMap* map16;
  const CUtensorMap* map1;
  const TileCursor& next2;
  int p;
  float2* slot;
  uint64_t* full;
  bool refill;
  uint64_t* exch_bar;
  uint32_t exch_parity;
  bool exch_wait;
  __device__ __forceinline__ void operator()() const {
    __syncwarp();
    if (refill && (threadIdx.x & 31) == 0) {
      fence_proxy_async();
      inv_issue_stage(P, map16, map1, next2, p, slot, full);
    }
    if (exch_wait) mbar_wait(exch_bar, exch_parity);
  }
};

template <int kFast>
__global__ void __launch_bounds__(kFwdThreads, 1) istft512_tma_kernel(const InvParams P, const __grid_constant__ CUtensorMap map16,
                                                                       const __grid_constant__ CUtensorMap map1) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2 * kInvBufs + 16];
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
  uint64_t* stage_full = bars + 2 * kInvBufs;      // [warp][slot], count 1 (+ tx bytes)
  float2* stage0 = reinterpret_cast<float2*>(smem);                         // 8 warps x 2 slots
  float2* exch0 = stage0 + 8 * 2 * kStageSlot;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kInvBufs; ++i) {
      mbar_init(exch_full + i, kThreads / 32);
      mbar_init(exch_empty + i, kThreads / 32);
    }
    for (int i = 0; i < 16; ++i) mbar_init(stage_full + i, 1);
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
    uint64_t* my_full = stage_full + 2 * warp;
    TileCursor c2 = c;                                  // cursor of the tile two ahead (the next refill)
    if (lane == 0) inv_issue_stage(P, &map16, &map1, c2, warp, my_stage, my_full);
    tile_advance(c2, P.tiles_per_clip);
    if (lane == 0 && n > 1) inv_issue_stage(P, &map16, &map1, c2, warp, my_stage + kStageSlot, my_full + 1);
    tile_advance(c2, P.tiles_per_clip);
    int es = 0, use = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const int s = i & 1;
      float2* slot = my_stage + s * kStageSlot;
      mbar_wait(my_full + s, (uint32_t)((i >> 1) & 1));
      InvLoadStaged load{slot + lane + ((c.tt * P.g.FO - P.g.HL) & 1), nullptr, nullptr};
      StagedBefore sb{P, &map16, &map1, c2, warp, slot, my_full + s, i + 2 < n,
                      exch_empty + es, (uint32_t)((use - 1) & 1), use >= 1};
      inv_stageA(exch0 + es * kExch, w, lane, warp, true, load, sb);
      mbar_arrive_warp(exch_full + es);
      tile_advance(c, P.tiles_per_clip);
      tile_advance(c2, P.tiles_per_clip);
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

// ---------------------------------------------------------------------------------------------------
// generic power-of-two path: one frame per CTA, radix-2 DIT in shared memory
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void smem_fft(float2* buf, int N, int logN, bool inverse) {
  // input already in bit-reversed order
  for (int s = 1; s <= logN; ++s) {
    const int half = 1 << (s - 1);
    for (int j = threadIdx.x; j < (N >> 1); j += blockDim.x) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << s) + pos;
      const int i1 = i0 + half;
      float sn, cs;
      sincospif((float)pos / (float)half, &sn, &cs);   // exp(-j pi pos/half) = cs - j sn
      if (inverse) sn = -sn;
      const float2 a = buf[i0], b = buf[i1];
      const float tr = b.x * cs + b.y * sn;
      const float ti = b.y * cs - b.x * sn;
      buf[i0] = make_float2(a.x + tr, a.y + ti);
      buf[i1] = make_float2(a.x - tr, a.y - ti);
    }
    __syncthreads();
  }
}

struct GenericFwdParams {
  FwdParams P;
  int N, logN, F;
};

__global__ void __launch_bounds__(256) stft_generic_fwd_kernel(const GenericFwdParams G) {
  extern __shared__ __align__(128) float smem[];
  float2* buf = reinterpret_cast<float2*>(smem);
  const FwdParams& P = G.P;
  const int N = G.N;
  for (int fix = blockIdx.x; fix < P.n_tiles; fix += gridDim.x) {
    const int b = (int)(fix / P.T_out);
    const int t = (int)(fix % P.T_out);
    const float* src = P.wave + (long long)b * P.wave_pitch;
    int gs = 0, ge = 0;
    if (P.gap_samples) { gs = P.gap_samples[2 * b]; ge = P.gap_samples[2 * b + 1]; }
    const long long g0 = (long long)t * P.hop - P.pad;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const long long g = g0 + n;
      float v = (g >= 0 && g < P.L && !(g >= gs && g < ge)) ? src[g] : 0.0f;
      v *= P.window[n];
      buf[__brev((unsigned)n) >> (32 - G.logN)] = make_float2(v, 0.0f);
    }
    __syncthreads();
    smem_fft(buf, N, G.logN, false);
    FwdEmitFull emit = fwd_make_emit_full(P, b, t, G.F, true);
    for (int k = threadIdx.x; k < G.F; k += blockDim.x) {
      const float2 x = buf[k];
      emit.rows(k, k);
      emit.put1(emit.lo(0), x.x, (k == 0 || k == N / 2) ? 0.0f : x.y);
    }
    __syncthreads();
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
  const long long total = (long long)P.B * P.out_len;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / P.out_len);
    const int s = (int)(i % P.out_len);
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

__global__ void __launch_bounds__(1024) db_heuristic_kernel(const float* x, long long n, int* flags) {
  __shared__ float smax[32];
  __shared__ double ssum[32];
  const float* xb = x + (long long)blockIdx.x * n;
  float mx = -INFINITY;
  double sm = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = xb[i];
    mx = fmaxf(mx, v);
    sm += (double)v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; ssum[threadIdx.x >> 5] = sm; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    mx = threadIdx.x < nw ? smax[threadIdx.x] : -INFINITY;
    sm = threadIdx.x < nw ? ssum[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
    }
    if (threadIdx.x == 0) flags[blockIdx.x] = (mx < 0.0f && sm < 0.0) ? 1 : 0;
  }
}

__global__ void gap_zero_kernel(const float* in, long long in_pitch, float* out, long long out_pitch,
                                long long B, long long L, const int* gaps) {
  const long long total = B * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / L, s = i % L;
    const bool in_gap = gaps && s >= gaps[2 * b] && s < gaps[2 * b + 1];
    out[b * out_pitch + s] = in_gap ? 0.0f : in[b * in_pitch + s];
  }
}

__global__ void gap_mask_kernel(float* mask, long long pitch, long long B, long long L, const int* gaps) {
  const long long total = B * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / L, s = i % L;
    const bool in_gap = gaps && s >= gaps[2 * b] && s < gaps[2 * b + 1];
    mask[b * pitch + s] = in_gap ? 0.0f : 1.0f;
  }
}

// one (clip, bin) row per WARP and loop trip: no per-element division, 16-byte stores where the row allows
__global__ void __launch_bounds__(256) frame_mask_kernel(float* mask, long long B, long long F, long long T, const int* fr,
                                                         int one_in_gap) {
  const long long rows = B * F;
  const float in_v = one_in_gap ? 1.0f : 0.0f, out_v = 1.0f - in_v;
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
    const long long b = row / F;
    const int f0 = fr ? fr[2 * b] : 0, f1 = fr ? fr[2 * b + 1] : 0;
    float* dst = mask + row * T;
    const int head = (int)((4 - ((reinterpret_cast<uintptr_t>(dst) >> 2) & 3)) & 3);       // elements before 16-byte alignment
    const int n4 = T > head ? (int)((T - head) >> 2) : 0;
    if (lane < head && lane < T) dst[lane] = (lane >= f0 && lane < f1) ? in_v : out_v;
    float4* d4 = reinterpret_cast<float4*>(dst + head);
    for (int q = lane; q < n4; q += 32) {
      const int t = head + 4 * q;
      d4[q] = make_float4((t >= f0 && t < f1) ? in_v : out_v, (t + 1 >= f0 && t + 1 < f1) ? in_v : out_v,
                          (t + 2 >= f0 && t + 2 < f1) ? in_v : out_v, (t + 3 >= f0 && t + 3 < f1) ? in_v : out_v);
    }
    const int t = head + 4 * n4 + lane;
    if (t < T) dst[t] = (t >= f0 && t < f1) ? in_v : out_v;
  }
}

// Gap variants: per-variant metadata for the transform kernel, {gap start, gap end, first re-transformed frame, wave row}
__global__ void variant_meta_kernel(const int* __restrict__ gaps, int4* __restrict__ meta, int B, int G, FwdParams P) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= B) return;
  const int gs = gaps[2 * v], ge = gaps[2 * v + 1];
  meta[v] = make_int4(gs, ge, var_frame_base(P, gs), v / G);
}

// Gap variants, copy pass: one (file i, bin k) row of the clean spectrogram per WARP and loop trip.  The row is read ONCE
// into registers (lanes along the frame axis, 16 x 32 frames per pass) and stored G times, into row k of each of the file's G
// variants -- a pure store stream like frame_mask_kernel.  Plain coalesced 4-byte accesses: source and destination rows sit at
// different 16-byte phases in general (T = 417, 834 are odd).  Frames [fb, fb + nt * kFR) of a variant are left to the
// transform kernel that runs next on the stream (var_frame_base); lane j of the warp holds fb of variant j.
// (Measured, 256 files x 25 gaps x 5 s: a row-per-variant copy that re-read the clean row for every variant ran at 2.9 TB/s,
// its dependent gap-start and L2 loads in front of every 1.7 KB row; 16-byte stores fed by 16-byte-strided scalar loads at 2.0.)
__global__ void __launch_bounds__(256) variant_fill_kernel(const float* __restrict__ clean, float* __restrict__ out, long long N,
                                                           int G, int F, int T, const int* __restrict__ gaps, FwdParams P) {
  const long long rows = N * F;
  const int lane = threadIdx.x & 31;
  const int span = P.tiles_per_clip * kFR;
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
    const long long i = row / F;
    const int k = (int)(row - i * F);
    const float* src = clean + row * T;
    for (int t0 = lane; t0 < T; t0 += 512) {
      float x[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = (t0 + 32 * j < T) ? __ldg(src + t0 + 32 * j) : 0.0f;
      for (int j0 = 0; j0 < G; j0 += 32) {
        const int mine = (j0 + lane < G) ? var_frame_base(P, gaps[2 * (i * G + j0 + lane)]) : 0;
        const int cnt = (G - j0) < 32 ? (G - j0) : 32;
        for (int jj = 0; jj < cnt; ++jj) {
          const int s0 = __shfl_sync(0xffffffffu, mine, jj), s1 = s0 + span;
          float* dst = out + (((i * G + j0 + jj) * F + k) * (long long)T);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int t = t0 + 32 * j;
            if (t < T && (t < s0 || t >= s1)) dst[t] = x[j];
          }
        }
      }
    }
  }
}

// Gap variants, copy pass through the TMA.  A variant's [F, T] block is one contiguous range, a copy of the file's clean block,
// so the copy is flat: a CTA stages a chunk of the clean block in shared memory and thread j sends it to variant j with ONE
// bulk copy shared -> global (UBLKCP).  Bulk copies need 16-byte aligned addresses on both sides, and the G destinations sit
// at up to four different 16-byte phases (F T = 257 * 417 is odd), so the chunk is staged FOUR times, copy h shifted by h
// elements (cph[m] = chunk[m + h]): the copy whose phase matches the destination feeds its aligned middle, the <= 3 + 3
// elements around it go by scalar stores.  No per-element store instructions (the scalar-store kernel above is limited by
// the LSU queue: ncu lg_throttle 16.7 stall cycles per issued instruction at 3.7 TB/s; a row-wise version of this kernel with
// one 1.6 KB bulk copy per row and variant reached 4.2 TB/s).  The frames the transform kernel rewrites afterwards are
// copied too.
constexpr int kFillChunk = 4096;        // floats per staged chunk: 4 x 16 KB of shared memory per CTA
constexpr int kFillThreads = 128;
__global__ void __launch_bounds__(kFillThreads) variant_fill_tma_kernel(const float* __restrict__ clean, float* __restrict__ out,
                                                                       long long N, int G, long long FT, int chunks_per_file) {
  extern __shared__ __align__(128) float vsm[];        // [4][kFillChunk]
  const int tid = threadIdx.x;
  const long long units = N * chunks_per_file;
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const long long i = u / chunks_per_file;
    const long long off = (u - i * chunks_per_file) * kFillChunk;
    const int len = (FT - off) < kFillChunk ? (int)(FT - off) : kFillChunk;
    const float* src = clean + i * FT + off;
    bulk_wait_read0();            // this thread's bulk copies of the previous chunk have read their source
    __syncthreads();
    for (int t0 = tid; t0 < len; t0 += 8 * kFillThreads) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (t0 + kFillThreads * j < len) ? __ldg(src + t0 + kFillThreads * j) : 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = t0 + kFillThreads * j;
        if (t < len) {
#pragma unroll
          for (int h = 0; h < 4; ++h)
            if (t >= h) vsm[h * kFillChunk + t - h] = x[j];
        }
      }
    }
    fence_proxy_async();          // the staged copies become visible to the async proxy
    __syncthreads();
    for (int j = tid; j < G; j += kFillThreads) {
      float* d = out + (i * G + j) * FT + off;
      int h = (int)((4 - ((reinterpret_cast<uintptr_t>(d) >> 2) & 3)) & 3);
      if (h > len) h = len;
      const int n4 = (len - h) >> 2;
      if (n4 > 0) tma_store_1d(d + h, vsm + h * kFillChunk, (uint32_t)n4 * 16u);
      bulk_commit();
      for (int e = 0; e < h; ++e) d[e] = vsm[e];
      for (int e = h + 4 * n4; e < len; ++e) d[e] = vsm[e];
    }
  }
  bulk_wait0();
}

__global__ void __launch_bounds__(256) peak_kernel(const float* in, long long pitch, long long L, float* peaks) {
  __shared__ float sm[8];
  const long long b = blockIdx.y;
  const float* x = in + b * pitch;
  float mx = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < L;
       i += (long long)gridDim.x * blockDim.x)
    mx = fmaxf(mx, fabsf(x[i]));
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sm[w]);
    // non-negative floats order like their bit patterns
    atomicMax(reinterpret_cast<int*>(peaks + b), __float_as_int(mx));
  }
}

__global__ void peak_scale_kernel(const float* in, long long in_pitch, float* out, long long out_pitch,
                                  long long B, long long L, const float* peaks) {
  const long long total = B * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / L, s = i % L;
    const float pk = peaks[b];
    const float v = in[b * in_pitch + s];
    out[b * out_pitch + s] = pk < kFltMin ? v : __fdiv_rn(v, pk);
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
    const float sa = m.x / (sqrtf(ax * ax + ay * ay) + kFltMin);
    const float sb = m.y / (sqrtf(bx * bx + by * by) + kFltMin);
    spec[i] = make_float4(ax * sa, ay * sa, bx * sb, by * sb);
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
    const float sa = m.x / (sqrtf(ax * ax + ay * ay) + kFltMin);
    const float sb = m.y / (sqrtf(bx * bx + by * by) + kFltMin);
    angles[i] = make_float4(ax * sa, ay * sa, bx * sb, by * sb);
  }
}

__global__ void gl_update_pp_tail_kernel(const float2* rebuilt, const float2* tprev, const float* mag, float2* angles,
                                         long long i, float alpha, int has_prev) {
  const float2 rb = rebuilt[i];
  float ax = rb.x, ay = rb.y;
  if (has_prev) { const float2 tp = tprev[i]; ax -= alpha * tp.x; ay -= alpha * tp.y; }
  const float s = mag[i] / (sqrtf(ax * ax + ay * ay) + kFltMin);
  angles[i] = make_float2(ax * s, ay * s);
}

__global__ void gl_update_tail_kernel(float2* spec, float2* tprev, const float* mag, long long i, float alpha, int has_prev) {
  const float2 rb = spec[i];
  float ax = rb.x, ay = rb.y;
  if (has_prev) { const float2 tp = tprev[i]; ax -= alpha * tp.x; ay -= alpha * tp.y; }
  tprev[i] = rb;
  const float s = mag[i] / (sqrtf(ax * ax + ay * ay) + kFltMin);
  spec[i] = make_float2(ax * s, ay * s);
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

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct DevInfo { int ok; int sms; int max_smem; };

static DevInfo dev_info() {
  DevInfo d{0, 0, 0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return d;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  d.ok = (major == 10);
  return d;
}

static inline int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
static inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
static inline int ew_grid(long long total, int sms) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)sms * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static long long num_frames(long long L, int n_fft, int hop, int center) {
  if (L < 0 || n_fft <= 0 || hop <= 0) return -1;
  const long long padded = L + (center ? 2LL * (n_fft / 2) : 0);
  if (padded < n_fft) return -1;
  return 1 + (padded - n_fft) / hop;
}

static long long istft_length(long long T, int n_fft, int hop, int center, long long length) {
  if (length > 0) return length;
  long long n = (long long)n_fft + (long long)hop * (T - 1);
  if (center) n -= 2LL * (n_fft / 2);
  return n;
}

static long long istft_used_frames(long long T, int n_fft, int hop, int center, long long length) {
  if (length <= 0) return T;
  const long long padded = length + (center ? 2LL * (n_fft / 2) : 0);
  const long long nf = (padded + hop - 1) / hop;
  return nf < T ? nf : T;
}

// shared memory of the forward kernel: n_tile_bufs staged-waveform buffers + 2 exchange buffers
static size_t fwd_smem_bytes(int hop, int n_tile_bufs) {
  return ((size_t)n_tile_bufs * (size_t)((fwd_tile_len(hop) + 31) & ~31) + 4 * (size_t)kExch) * sizeof(float);
}

static int fwd_tile_bufs(const aip_stft_desc* d, const DevInfo& di) {
  if (d->n_fft != 512 || (d->hop & 1)) return 0;
  int most = kFwdTileBufs;
  if (const char* e = getenv("AIP_FWD_TILE_BUFS")) { const int v = atoi(e); if (v >= 1 && v <= kFwdTileBufs) most = v; }   // profiling switch
  for (int nb = most; nb >= 1; --nb)
    if (fwd_smem_bytes(d->hop, nb) + 6 * 1024 <= (size_t)di.max_smem) return nb;       // + static tables and barriers
  return 0;
}

static bool fwd_fast_ok(const aip_stft_desc* d, const DevInfo& di) { return fwd_tile_bufs(d, di) > 0; }

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
  // Off unless AIP_INV_TMA=1: measured 0.568 ms against 0.535 ms for the direct-load kernel (1024 x 10 s, hop 192) -- the
  // staging slots leave room for ONE exchange buffer only, and stage A is not the slower role (profiles/README.md).
  const char* e = getenv("AIP_INV_TMA");
  return e && atoi(e) != 0;
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

// Tile counters of the dynamic schedule: one 4-byte slot per launch, zeroed on the launch's stream right before it.
// 64 slots rotate, so launches stay independent unless more than 64 of them are in flight at once.
__device__ unsigned g_tile_counters[64];
static unsigned* next_tile_counter(cudaStream_t st, cudaError_t* err) {
  static std::atomic<unsigned> launch_id{0};
  unsigned* base = nullptr;
  *err = cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_tile_counters);
  if (*err != cudaSuccess) return nullptr;
  unsigned* slot = base + (launch_id.fetch_add(1, std::memory_order_relaxed) & 63u);
  *err = cudaMemsetAsync(slot, 0, sizeof(unsigned), st);
  return slot;
}

template <int kMode>
static cudaError_t launch_fwd512_t(FwdParams P, const DevInfo& di, cudaStream_t st) {
  auto kern = P.zero_groups == 2 ? stft512_fwd_kernel<kMode, 2> : stft512_fwd_kernel<kMode, 0>;
  // shape-specialised builds of the log-magnitude variant for the reference's fixed shapes (config.py: n_fft 512 / win 384 /
  // hop 192; 5 s clips -> 417 frames, models/CNNBLSTM/dataset.py:89; 10 s -> 834): store offsets become immediates
  if (kMode == FWD_MAG_LOG10 && P.zero_groups == 2 && !getenv("AIP_FWD_NO_SHAPE")) {
    if (P.T_out == 834) kern = stft512_fwd_kernel<FWD_MAG_LOG10, 2, 834>;
    else if (P.T_out == 417) kern = stft512_fwd_kernel<FWD_MAG_LOG10, 2, 417>;
  }
  const size_t smem = fwd_smem_bytes(P.hop, P.n_tile_bufs);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int grid = di.sms;                    // persistent: one CTA per SM
  // Tiles per draw.  Besides balancing the SMs, the chunk size sets how close in memory the tiles are that the 148 CTAs
  // work on at the same time: measured 1.73 / 1.58 / 1.52 / 1.50 / 1.51 / 1.55 / 1.58 ms for 1 / 4 / 8 / 12 / 16 / 32 / 64
  // tiles per draw on the log-magnitude variant (4096 x 10 s; static equal runs: 1.60 ms), while the complex-output variant
  // (8 bytes per bin) is fastest at 2 (0.468 ms against 0.529 ms static, 0.507 ms at 8).
  const int out_bytes = (P.mag ? 4 : 0) + (P.phase ? 4 : 0) + (P.mask ? 4 : 0) + (P.spec ? 8 : 0);
  P.chunk = out_bytes <= 4 ? 12 : (P.spec ? 2 : 4);
  if (kMode & FWD_VARIANT) P.chunk = 4;      // few tiles per CTA (one or two per variant): balance matters more than locality
  if (const char* c = getenv("AIP_FWD_CHUNK")) { const int cv = atoi(c); if (cv >= 1 && cv <= 4096) P.chunk = cv; }    // profiling switch
  const int n_chunks = (P.n_tiles + P.chunk - 1) / P.chunk;
  if (grid > n_chunks) grid = n_chunks;
  P.tile_counter = next_tile_counter(st, &e);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)grid, kFwdThreads, smem, st>>>(P);
  return cudaGetLastError();
}

static cudaError_t launch_fwd512(const FwdParams& P, const DevInfo& di, cudaStream_t st) {
  if (P.var_div > 0) {      // gap variants: magnitude-only epilogues
    switch (fwd_mode_of(P)) {
      case FWD_MAG_ABS: return launch_fwd512_t<FWD_MAG_ABS | FWD_VARIANT>(P, di, st);
      case FWD_MAG_LOG10: return launch_fwd512_t<FWD_MAG_LOG10 | FWD_VARIANT>(P, di, st);
      case MAG_LOG1P_POW: return launch_fwd512_t<MAG_LOG1P_POW | FWD_VARIANT>(P, di, st);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (fwd_mode_of(P)) {
    case FWD_MAG_ABS: return launch_fwd512_t<FWD_MAG_ABS>(P, di, st);
    case FWD_MAG_LOG10: return launch_fwd512_t<FWD_MAG_LOG10>(P, di, st);
    case FWD_SPEC: return launch_fwd512_t<FWD_SPEC>(P, di, st);
    case MAG_LOG10_EPS | FWD_MASK: return launch_fwd512_t<MAG_LOG10_EPS | FWD_MASK>(P, di, st);
    case MAG_LOG1P_POW: return launch_fwd512_t<MAG_LOG1P_POW>(P, di, st);
    case MAG_LOG1P_POW | FWD_PHASE | FWD_MASK: return launch_fwd512_t<MAG_LOG1P_POW | FWD_PHASE | FWD_MASK>(P, di, st);
    case FWD_SPEC | FWD_PHASE | FWD_MASK: return launch_fwd512_t<FWD_SPEC | FWD_PHASE | FWD_MASK>(P, di, st);
    case MAG_LOG10_EPS | FWD_ZERO: return launch_fwd512_t<MAG_LOG10_EPS | FWD_ZERO>(P, di, st);
    case MAG_ABS | FWD_PHASE: return launch_fwd512_t<MAG_ABS | FWD_PHASE>(P, di, st);
    case MAG_LOG1P_POW | FWD_PHASE: return launch_fwd512_t<MAG_LOG1P_POW | FWD_PHASE>(P, di, st);
    case FWD_SPEC | FWD_PHASE: return launch_fwd512_t<FWD_SPEC | FWD_PHASE>(P, di, st);
    default: return launch_fwd512_t<FWD_FULL>(P, di, st);
  }
}

static cudaError_t launch_fwd_generic(FwdParams P, int n_fft, const DevInfo& di, cudaStream_t st) {
  GenericFwdParams G;
  G.N = n_fft; G.logN = ilog2(n_fft); G.F = n_fft / 2 + 1;
  if ((long long)P.B * P.T_out > 0x7fffffffLL) return cudaErrorInvalidValue;
  P.n_tiles = (int)((long long)P.B * P.T_out);
  G.P = P;
  const size_t smem = (size_t)n_fft * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(stft_generic_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long grid = (long long)di.sms * 8;
  if (grid > P.n_tiles) grid = P.n_tiles;
  stft_generic_fwd_kernel<<<(unsigned)grid, 256, smem, st>>>(G);
  return cudaGetLastError();
}

// fills the geometry-dependent fields and launches; P carries pointers / epilogue settings already
static int run_fwd(const aip_stft_desc* desc, FwdParams P, long long T_out, cudaStream_t st) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!desc || !desc->window || !P.wave) return AIP_ERR_ARG;
  if (!is_pow2(desc->n_fft) || desc->n_fft < 32 || desc->n_fft > 4096 || desc->hop <= 0) return AIP_ERR_UNSUPPORTED;
  if (P.B < 0 || P.L < 0 || P.wave_pitch < P.L) return AIP_ERR_ARG;
  const long long T = num_frames(P.L, desc->n_fft, desc->hop, desc->center);
  if (T < 1 || T_out < 0 || T_out > T) return AIP_ERR_ARG;
  if ((P.mag_kind != MAG_NONE) != (P.mag != nullptr)) return AIP_ERR_ARG;
  if (P.mag_kind < MAG_NONE || P.mag_kind > MAG_POW) return AIP_ERR_ARG;
  if (P.B == 0 || T_out == 0) return AIP_OK;
  P.hop = desc->hop;
  P.pad = desc->center ? desc->n_fft / 2 : 0;
  P.T = (int)T;
  P.T_out = (int)T_out;
  P.window = desc->window;
  cudaError_t e;
  if (fwd_fast_ok(desc, di)) {
    P.tiles_per_clip = (int)((T_out + kFR - 1) / kFR);
    // (the dynamic schedule's counter overshoots n_tiles by up to grid * chunk draws: keep clear of the int range)
    if ((long long)P.B * P.tiles_per_clip > 0x3fffffffLL || T_out > (1 << 22)) return AIP_ERR_UNSUPPORTED;
    P.n_tiles = (int)((long long)P.B * P.tiles_per_clip);
    P.tile_floats = (fwd_tile_len(P.hop) + 31) & ~31;
    P.n_tile_bufs = fwd_tile_bufs(desc, di);
    // three buffers (two bulk copies in flight) measured faster than two for every variant: 1.584 -> 1.503 ms (log-magnitude),
    // 0.469 -> 0.458 ms (complex output), both under the dynamic tile schedule
    P.zero_groups = win_zero_groups(desc->win_length);
    P.vec_ok = ((P.hop & 3) == 0) && ((P.pad & 3) == 0) && ((P.wave_pitch & 3) == 0) &&
               ((reinterpret_cast<uintptr_t>(P.wave) & 15) == 0);
    e = launch_fwd512(P, di, st);
  } else {
    e = launch_fwd_generic(P, desc->n_fft, di, st);
  }
  return (int)e;
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
    P.ola_fast = inv_ola_fast_kind(P.hop, P.pad, desc->win_length);
    if (const char* of = getenv("AIP_OLA_FAST")) { if (!((atoi(of) >> (P.ola_fast - 1)) & 1)) P.ola_fast = 0; }   // profiling switch (bit mask of kinds)
    P.n_bufs = kInvBufsDefault;
    if (const char* nb = getenv("AIP_INV_BUFS")) { const int v = atoi(nb); if (v >= 1 && v <= kInvBufs) P.n_bufs = v; }   // profiling switch
    int grid = di.sms;
    if (grid > P.n_tiles) grid = P.n_tiles;
    P.tiles_per_cta = (P.n_tiles + grid - 1) / grid;
    grid = (P.n_tiles + P.tiles_per_cta - 1) / P.tiles_per_cta;
    if (inv_tma_ok(P)) {
      CUtensorMap map16, map1;
      if (inv_make_map(&map16, P.spec, P.B, P.T, P.n_frames, 16) && inv_make_map(&map1, P.spec, P.B, P.T, P.n_frames, 1)) {
        P.n_bufs = 1;      // 8 x 2 staging slots (142 KB) + one exchange buffer (66 KB)
        const size_t smem_tma = (size_t)8 * kStageBytesWarp + (size_t)kExch * sizeof(float2);
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
  } else {
    const size_t need = aip_istft_workspace_bytes(desc, P.B, P.T);
    if (!workspace || workspace_bytes < need) return AIP_ERR_WORKSPACE;
    GenericInvParams G;
    G.N = desc->n_fft; G.logN = ilog2(desc->n_fft); G.F = desc->n_fft / 2 + 1;
    G.frames = static_cast<float*>(workspace);
    if ((long long)P.B * P.n_frames > 0x7fffffffLL) return AIP_ERR_UNSUPPORTED;
    P.n_tiles = (int)((long long)P.B * P.n_frames);
    G.P = P;
    const size_t smem = (size_t)desc->n_fft * sizeof(float2);
    e = cudaFuncSetAttribute(istft_generic_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    long long grid = (long long)di.sms * 8;
    if (grid > P.n_tiles) grid = P.n_tiles;
    istft_generic_frames_kernel<<<(unsigned)grid, 256, smem, st>>>(G);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    istft_generic_ola_kernel<<<ew_grid((long long)P.B * P.out_len, di.sms), 256, 0, st>>>(G);
    e = cudaGetLastError();
  }
  return (int)e;
}

}  // namespace aip

// ===================================================================================================
// C ABI
// ===================================================================================================
using namespace aip;

extern "C" {

int64_t aip_num_frames(int64_t L, int32_t n_fft, int32_t hop, int32_t center) {
  return num_frames(L, n_fft, hop, center);
}

int64_t aip_istft_length(int64_t T, int32_t n_fft, int32_t hop, int32_t center, int64_t length) {
  if (T < 1 || n_fft <= 0 || hop <= 0 || length < 0) return -1;
  return istft_length(T, n_fft, hop, center, length);
}

int aip_stft_fwd_f32(const aip_stft_desc* desc, const float* wave, int64_t B, int64_t L, int64_t wave_pitch,
                     const int32_t* gap_samples, const int32_t* zero_frames, const int32_t* mask_frames,
                     int32_t mask_in_gap_is_one, int32_t mag_kind, float eps, float power, int64_t T_out,
                     float* spec_out, float* mag_out, float* phase_out, float* mask_out, void* stream) {
  if (B > 0x7fffffffLL || L > 0x7fffffffLL) return AIP_ERR_ARG;
  FwdParams P{};
  P.wave = wave; P.wave_pitch = wave_pitch; P.B = (int)B; P.L = (int)L;
  P.gap_samples = gap_samples; P.zero_frames = zero_frames; P.mask_frames = mask_frames;
  P.mask_in_gap_is_one = mask_in_gap_is_one;
  P.mag_kind = mag_kind; P.eps = eps; P.power = power;
  P.spec = reinterpret_cast<float2*>(spec_out); P.mag = mag_out; P.phase = phase_out; P.mask = mask_out;
  return run_fwd(desc, P, T_out, static_cast<cudaStream_t>(stream));
}

int aip_istft_blend_f32(const aip_stft_desc* desc, const float* model_out, const float* blend_in, const float* blend_mask,
                        const float* phase, int32_t mag_domain, int64_t B, int64_t T, int64_t length,
                        const float* inv_wss, float* wave_out, int64_t out_pitch, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (B > 0x7fffffffLL || T > 0x7fffffffLL) return AIP_ERR_ARG;
  if (!model_out || !blend_in || !blend_mask || !phase) return AIP_ERR_ARG;
  if (mag_domain != DOM_POW10 && mag_domain != DOM_DB) return AIP_ERR_UNSUPPORTED;
  InvParams P{};
  P.mag = model_out; P.blend_in = blend_in; P.blend_mask = blend_mask; P.phase = phase; P.mag_domain = mag_domain;
  P.B = (int)B; P.T = (int)T; P.inv_wss = inv_wss; P.out = wave_out; P.out_pitch = out_pitch;
  return run_inv(desc, P, length, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t aip_istft_workspace_bytes(const aip_stft_desc* desc, int64_t B, int64_t T) {
  if (!desc || B <= 0 || T <= 0 || desc->n_fft <= 0 || inv_fast_ok(desc)) return 0;
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
                             const float* inv_wss, float* wave_out, int64_t out_pitch, float* peaks, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (B > 0x7fffffffLL || T > 0x7fffffffLL || B < 0 || !peaks || !desc) return AIP_ERR_ARG;
  if (B == 0) return AIP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  cudaError_t e = cudaMemsetAsync(peaks, 0, (size_t)B * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const bool fused = inv_fast_ok(desc);
  InvParams P{};
  P.spec = reinterpret_cast<const float2*>(spec); P.mag = mag; P.phase = phase; P.mag_domain = mag_domain;
  P.db_flags = db_flags; P.B = (int)B; P.T = (int)T; P.inv_wss = inv_wss; P.out = wave_out; P.out_pitch = out_pitch;
  P.peaks = fused ? peaks : nullptr;
  const int rc = run_inv(desc, P, length, workspace, workspace_bytes, st);
  if (rc != AIP_OK) return rc;
  const long long out_len = istft_length(T, desc->n_fft, desc->hop, desc->center, length);
  if (out_len <= 0) return AIP_OK;
  if (!fused) {      // generic n_fft path: separate peak pass
    if (B > 65535) return AIP_ERR_UNSUPPORTED;
    long long gx = (out_len + 256 * 8 - 1) / (256 * 8);
    if (gx > 64) gx = 64;
    peak_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, st>>>(wave_out, out_pitch, out_len, peaks);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  peak_scale_kernel<<<ew_grid(B * out_len, di.sms), 256, 0, st>>>(wave_out, out_pitch, wave_out, out_pitch, B, out_len, peaks);
  return (int)cudaGetLastError();
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
  const bool fused = pingpong && fwd_fast_ok(desc, di) && inv_fast_ok(desc) && !getenv("AIP_GL_UNFUSED");
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

int aip_db_heuristic_f32(const float* x, int64_t B, int64_t n, int32_t* flags, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!x || !flags || B < 0 || n < 1 || B > 0x7fffffffLL) return AIP_ERR_ARG;
  if (B == 0) return AIP_OK;
  db_heuristic_kernel<<<(unsigned)B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(x, n, flags);
  return (int)cudaGetLastError();
}

int aip_gap_zero_f32(const float* in, int64_t in_pitch, float* out, int64_t out_pitch, int64_t B, int64_t L,
                     const int32_t* gap_samples, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!in || !out || B < 0 || L < 0 || in_pitch < L || out_pitch < L) return AIP_ERR_ARG;
  if (B * L == 0) return AIP_OK;
  gap_zero_kernel<<<ew_grid(B * L, di.sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, in_pitch, out, out_pitch, B, L, gap_samples);
  return (int)cudaGetLastError();
}

int aip_gap_mask_f32(float* mask, int64_t pitch, int64_t B, int64_t L, const int32_t* gap_samples, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!mask || B < 0 || L < 0 || pitch < L) return AIP_ERR_ARG;
  if (B * L == 0) return AIP_OK;
  gap_mask_kernel<<<ew_grid(B * L, di.sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, pitch, B, L, gap_samples);
  return (int)cudaGetLastError();
}

int aip_frame_mask_f32(float* mask, int64_t B, int64_t F, int64_t T, const int32_t* mask_frames,
                       int32_t mask_in_gap_is_one, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!mask || B < 0 || F < 0 || T < 0) return AIP_ERR_ARG;
  if (B * F * T == 0) return AIP_OK;
  long long grid = (B * F + 7) / 8;
  if (grid > (long long)di.sms * 16) grid = (long long)di.sms * 16;
  frame_mask_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, B, F, T, mask_frames, mask_in_gap_is_one);
  return (int)cudaGetLastError();
}

size_t aip_stft_gap_variants_workspace_bytes(int64_t N, int64_t G) {
  return (N > 0 && G > 0) ? (size_t)N * (size_t)G * 16u : 0;
}

int aip_stft_gap_variants_f32(const aip_stft_desc* desc, const float* wave, int64_t N, int64_t L, int64_t wave_pitch,
                              int64_t G, const int32_t* gap_samples, int32_t gap_len_max, int32_t mag_kind, float eps,
                              int64_t T_out, const float* clean_mag, float* mag_out, void* workspace, size_t workspace_bytes,
                              void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!desc || !desc->window || !wave || !gap_samples || !clean_mag || !mag_out) return AIP_ERR_ARG;
  if (N < 0 || G < 1 || L < 0 || wave_pitch < L || gap_len_max < 0 || N * G > 0x7fffffffLL || L > 0x7fffffffLL) return AIP_ERR_ARG;
  if (mag_kind != MAG_ABS && mag_kind != MAG_LOG10_EPS && mag_kind != MAG_LOG1P_POW) return AIP_ERR_UNSUPPORTED;
  if (!fwd_fast_ok(desc, di)) return AIP_ERR_UNSUPPORTED;       // n_fft 512 register-FFT path only
  const long long T = num_frames(L, desc->n_fft, desc->hop, desc->center);
  if (T < 1 || T_out < 0 || T_out > T || T_out > (1 << 22)) return AIP_ERR_ARG;
  if (N == 0 || T_out == 0) return AIP_OK;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 15)) return AIP_ERR_ARG;
  if (workspace_bytes < aip_stft_gap_variants_workspace_bytes(N, G)) return AIP_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FwdParams P{};
  P.wave = wave; P.wave_pitch = wave_pitch; P.B = (int)(N * G); P.L = (int)L;
  P.gap_samples = gap_samples; P.mag_kind = mag_kind; P.eps = eps; P.power = 1.0f; P.mag = mag_out;
  P.hop = desc->hop; P.pad = desc->center ? desc->n_fft / 2 : 0;
  P.T = (int)T; P.T_out = (int)T_out; P.window = desc->window;
  P.var_div = (int)G;
  P.tiles_per_clip = var_tiles(gap_len_max, P.hop, P.T_out);
  if ((long long)P.B * P.tiles_per_clip > 0x3fffffffLL) return AIP_ERR_UNSUPPORTED;
  P.n_tiles = P.B * P.tiles_per_clip;
  P.tile_floats = (fwd_tile_len(P.hop) + 31) & ~31;
  P.n_tile_bufs = fwd_tile_bufs(desc, di);
  P.zero_groups = win_zero_groups(desc->win_length);
  P.vec_ok = ((P.hop & 3) == 0) && ((P.pad & 3) == 0) && ((P.wave_pitch & 3) == 0) &&
             ((reinterpret_cast<uintptr_t>(P.wave) & 15) == 0);
  const int F = desc->n_fft / 2 + 1;
  cudaError_t e;
  P.var_meta = static_cast<const int*>(workspace);
  P.var_prefetch = getenv("AIP_VAR_NO_PREFETCH") ? 0 : 1;                 // profiling switch
  variant_meta_kernel<<<(unsigned)((P.B + 255) / 256), 256, 0, st>>>(gap_samples, static_cast<int4*>(workspace), P.B, (int)G, P);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const size_t fill_smem = (size_t)4 * kFillChunk * sizeof(float);
  const char* fill_env = getenv("AIP_VAR_FILL");                       // profiling switch: "scalar" = the store-instruction kernel
  const bool aligned4 = ((reinterpret_cast<uintptr_t>(mag_out) | reinterpret_cast<uintptr_t>(clean_mag)) & 3) == 0;
  if (getenv("AIP_VAR_NO_FILL")) {
    // profiling switch: transform kernel only (tools/variant_probe.py)
  } else if (aligned4 && !(fill_env && fill_env[0] == 's')) {
    e = cudaFuncSetAttribute(variant_fill_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem);
    if (e != cudaSuccess) return (int)e;
    const long long FT = (long long)F * P.T_out;
    const long long cpf = (FT + kFillChunk - 1) / kFillChunk;
    if (cpf > 0x7fffffffLL) return AIP_ERR_UNSUPPORTED;
    long long grid = (long long)N * cpf;
    if (grid > (long long)di.sms * 3) grid = (long long)di.sms * 3;            // 3 x 64 KB of staging per SM
    variant_fill_tma_kernel<<<(unsigned)grid, kFillThreads, fill_smem, st>>>(clean_mag, mag_out, (long long)N, (int)G, FT, (int)cpf);
  } else {
    long long grid = ((long long)N * F + 7) / 8;
    if (grid > (long long)di.sms * 8) grid = (long long)di.sms * 8;
    variant_fill_kernel<<<(unsigned)grid, 256, 0, st>>>(clean_mag, mag_out, (long long)N, (int)G, F, P.T_out, gap_samples, P);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return (int)launch_fwd512(P, di, st);
}

int aip_peak_normalize_f32(const float* in, int64_t in_pitch, float* out, int64_t out_pitch, int64_t B,
                           int64_t L, float* peaks, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!in || !out || !peaks || B < 0 || L < 0 || in_pitch < L || out_pitch < L || B > 65535) return AIP_ERR_ARG;
  if (B * L == 0) return AIP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(peaks, 0, (size_t)B * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  long long gx = (L + 256 * 8 - 1) / (256 * 8);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  peak_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, st>>>(in, in_pitch, L, peaks);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  peak_scale_kernel<<<ew_grid(B * L, di.sms), 256, 0, st>>>(in, in_pitch, out, out_pitch, B, L, peaks);
  return (int)cudaGetLastError();
}

const char* aip_status_string(int status) {
  switch (status) {
    case AIP_OK: return "ok";
    case AIP_ERR_ARG: return "invalid argument";
    case AIP_ERR_UNSUPPORTED: return "unsupported parameter combination";
    case AIP_ERR_DEVICE: return "current CUDA device is not sm_100 (B200); there is no fallback path";
    case AIP_ERR_WORKSPACE: return "workspace missing or too small";
    default: return status > 0 ? cudaGetErrorString(static_cast<cudaError_t>(status)) : "unknown status";
  }
}

const char* aip_version(void) { return "aip_b200 0.1.0 sm_100a"; }

int aip_device_supported(void) { return dev_info().ok; }

}  // extern "C"
