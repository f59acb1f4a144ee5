// Forward transform: sm_100a kernels + their launchers + the forward entry points of the C ABI (include/aip_b200.h).
//
//   stft512_fwd_kernel   persistent; tile = 32 frames of one clip; waveform staged once in shared
//                        memory by TMA bulk copies, register 16x16 FFT (zero window taps pruned), packed
//                        split pass, fused |S| / log / phase / mask epilogues, coalesced [F,T] stores.
//   gap variants         (aip_stft_gap_variants_f32: G gapped spectrograms per file from ONE clean transform)
//                        variant_meta_kernel -> variant_fill_tma_kernel (clean block staged in shared memory at the four
//                        16-byte phases, one bulk copy shared -> global per chunk and variant) -> stft512_fwd_kernel
//                        <mode | FWD_VARIANT> (re-transform of the one or two tiles a gap touches).
//   stft_generic_fwd_kernel   any power-of-two n_fft in [32, 4096] (or odd hop): one frame per CTA, shared-memory
//                        radix-2.  Correct, not tuned: the reference's models only ever use n_fft = 512
//                        (config.py:28, GAN/config.yaml:12).
#include <mutex>
#include <unordered_map>

#include "aip_device.cuh"
#include "aip_host.h"

namespace aip {

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

struct WaitEmpty {
  uint64_t* bar;
  uint32_t parity;
  bool enabled;
  __device__ __forceinline__ void operator()() const { if (enabled) mbar_wait(bar, parity); }
};

struct ArriveRelease {
  uint64_t* bar;
  __device__ __forceinline__ void operator()() const { mbar_arrive_warp(bar); }
};

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
    int next_s = 0;                          // the next chunk, drawn ahead
    bool have_next = false;
    auto request = [&](float* buf, uint64_t* bar) {      // ptid == 0: start the copy of the next tile, if there is one
      if (fn.left == 0) {
        const int s = have_next ? next_s : (int)atomicAdd(P.tile_counter, (unsigned)P.chunk);
        have_next = false;
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
      // The next draw (an atomic on L2, ~1 us under load) is issued now and read one request later: 0.464 -> 0.455 ms (hop 192),
      // 0.709 -> 0.679 ms (hop 128) for 1024 clips with complex output, where the stage-1 warps are the slower role.  (Where
      // stage 2 is the slower role it cost 1 % as long as stage 2 handed the exchange buffer back early -- see fwd_phase2 --
      // and is neutral since.)
      if (fn.left == 0) {
        next_s = (int)atomicAdd(P.tile_counter, (unsigned)P.chunk);
        have_next = true;
      }
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
      // Stage 1 loads, windows and transforms the tile BEFORE it waits for the exchange buffer (its results sit in registers
      // anyway); only the twiddle / store loop runs after the hand-off.  With stage 2's late release (fwd_phase2) the FP32-heavy
      // part of stage 1 then runs beside stage 2's last epilogue rounds and its shared-memory loads of the next tile, and the tile
      // is delivered 0.4 instead of 1.0 stage-1 phases after the release: headline 1.410 -> 1.388 ms with the release moved from
      // 3 to 4 rounds (alone, at 3 rounds: 1.407; 5 / 6 / 7 / 8 rounds: 1.387 - 1.41 / 1.417 / 1.44 - 1.49 / 1.48 - 1.50).
      WaitEmpty we{exch_empty + es, (uint32_t)(((i >> 1) - 1) & 1), i >= 2};
      fwd_phase1<kZP>(P, ptid, tile, exch0 + es * kExch, win_s, lc, we);
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
      long long g = g0 + n;
      if (P.reflect && (g < 0 || g >= P.L)) g = g < 0 ? -g : 2LL * (P.L - 1) - g;
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

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// shared memory of the forward kernel: n_tile_bufs staged-waveform buffers + 2 exchange buffers
static size_t fwd_smem_bytes(int hop, int n_tile_bufs) {
  return ((size_t)n_tile_bufs * (size_t)((fwd_tile_len(hop) + 31) & ~31) + 4 * (size_t)kExch) * sizeof(float);
}

static int fwd_tile_bufs(const aip_stft_desc* d, const DevInfo& di) {
  if (d->n_fft != 512 || (d->hop & 1)) return 0;
  int most = kFwdTileBufs;
  if (const int v = tunables().fwd_tile_bufs) { if (v >= 1 && v <= kFwdTileBufs) most = v; }
  for (int nb = most; nb >= 1; --nb)
    if (fwd_smem_bytes(d->hop, nb) + 6 * 1024 <= (size_t)di.max_smem) return nb;       // + static tables and barriers
  return 0;
}

bool fwd_fast_ok(const aip_stft_desc* d, const DevInfo& di) { return fwd_tile_bufs(d, di) > 0; }

// Tile counters of the dynamic schedule.  A counter must never be shared by two launches that can be in flight together.
// Launches on ONE stream run in order, so each (device, stream) pair owns one 4-byte slot of g_tile_counters for as long as
// the library lives; the slot is zeroed by a memset enqueued on that stream right before the kernel, and the memset + launch
// pair is enqueued under a lock, so host threads that share a stream cannot interleave their pairs.  Launches on different
// streams (or devices: the symbol has one instance per device) use different slots and are independent however many are in
// flight.  kCounterSlots streams per device are supported; a further stream returns cudaErrorLaunchOutOfResources.
constexpr int kCounterSlots = 4096;
__device__ unsigned g_tile_counters[kCounterSlots];

struct CounterPool {
  std::mutex mu;
  std::unordered_map<unsigned long long, int> slot_of;      // (device << 48) ^ stream handle -> slot
  int used[64] = {0};                                       // slots handed out per device
};
static CounterPool& counter_pool() { static CounterPool p; return p; }

// enqueues "counter := 0" and then `launch(counter)` on `st`, atomically with respect to other host threads
template <class Launch>
static cudaError_t with_tile_counter(cudaStream_t st, Launch launch) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  unsigned* base = nullptr;
  e = cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_tile_counters);
  if (e != cudaSuccess) return e;
  CounterPool& pool = counter_pool();
  std::lock_guard<std::mutex> lock(pool.mu);
  const unsigned long long key = ((unsigned long long)dev << 48) ^ (unsigned long long)reinterpret_cast<uintptr_t>(st);
  auto it = pool.slot_of.find(key);
  if (it == pool.slot_of.end()) {
    if (pool.used[dev] >= kCounterSlots) return cudaErrorLaunchOutOfResources;
    it = pool.slot_of.emplace(key, pool.used[dev]++).first;
  }
  unsigned* slot = base + it->second;
  e = cudaMemsetAsync(slot, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  return launch(slot);
}

template <int kMode>
static cudaError_t launch_fwd512_t(FwdParams P, const DevInfo& di, cudaStream_t st) {
  auto kern = P.zero_groups == 2 ? stft512_fwd_kernel<kMode, 2> : stft512_fwd_kernel<kMode, 0>;
  // shape-specialised builds of the log-magnitude variant for the reference's fixed shapes (config.py: n_fft 512 / win 384 /
  // hop 192; 5 s clips -> 417 frames, models/CNNBLSTM/dataset.py:89; 10 s -> 834): store offsets become immediates
  if (kMode == FWD_MAG_LOG10 && P.zero_groups == 2 && !tunables().fwd_no_shape) {
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
  // (8 bytes per bin) is fastest at 2 (0.468 ms against 0.529 ms static, 0.507 ms at 8).  Re-measured after the late release of
  // the exchange buffer (fwd_phase2) and the draw-ahead: 1.446 / 1.424 / 1.415 / 1.411 / 1.413 / 1.416 / 1.429 / 1.480 ms for
  // 4 / 6 / 7 / 8 / 9 / 10 / 12 / 16 tiles per draw (hop 128, 1024 clips: 0.587 ms at 7 - 8, 0.604 at 12).
  const int out_bytes = (P.mag ? 4 : 0) + (P.phase ? 4 : 0) + (P.mask ? 4 : 0) + (P.spec ? 8 : 0);
  P.chunk = out_bytes <= 4 ? 8 : (P.spec ? 2 : 4);
  if (kMode & FWD_VARIANT) P.chunk = 4;      // few tiles per CTA (one or two per variant): balance matters more than locality
  if (const int cv = tunables().fwd_chunk) { if (cv >= 1 && cv <= 4096) P.chunk = cv; }
  const int n_chunks = (P.n_tiles + P.chunk - 1) / P.chunk;
  if (grid > n_chunks) grid = n_chunks;
  return with_tile_counter(st, [&](unsigned* counter) {
    P.tile_counter = counter;
    kern<<<(unsigned)grid, kFwdThreads, smem, st>>>(P);
    return cudaGetLastError();
  });
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
    case MAG_POW: return launch_fwd512_t<MAG_POW>(P, di, st);
    case FWD_SPEC | MAG_ABS: return launch_fwd512_t<FWD_SPEC | MAG_ABS>(P, di, st);
    case FWD_SPEC | MAG_LOG10_EPS: return launch_fwd512_t<FWD_SPEC | MAG_LOG10_EPS>(P, di, st);
    case FWD_SPEC | MAG_LOG1P_POW: return launch_fwd512_t<FWD_SPEC | MAG_LOG1P_POW>(P, di, st);
    case MAG_LOG10_EPS | FWD_PHASE: return launch_fwd512_t<MAG_LOG10_EPS | FWD_PHASE>(P, di, st);
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
  if ((long long)P.B * P.T_out > 0x7fffffffLL || (long long)G.F * P.T_out > 0x7fffffffLL) return cudaErrorInvalidValue;
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
int run_fwd(const aip_stft_desc* desc, FwdParams P, long long T_out, cudaStream_t st) {
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
  if (!P.spec && !P.mag && !P.phase && !P.mask) return AIP_OK;      // no output requested: nothing to launch
  P.hop = desc->hop;
  P.pad = desc->center ? desc->n_fft / 2 : 0;
  P.reflect = desc->center == 2;
  if (P.reflect && P.L <= P.pad) return AIP_ERR_UNSUPPORTED;        // np.pad(mode="reflect") needs pad < L
  P.T = (int)T;
  P.T_out = (int)T_out;
  P.window = desc->window;
  cudaError_t e;
  if (fwd_fast_ok(desc, di)) {
    P.tiles_per_clip = (int)((T_out + kFR - 1) / kFR);
    // (the dynamic schedule's counter overshoots n_tiles by up to 2 * grid * chunk -- every CTA draws one chunk ahead: keep clear of the int range)
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
  } else if (pow2_ok(desc->n_fft)) {
    e = launch_fwd_pow2(P, desc->n_fft, di, st);
  } else {
    e = launch_fwd_generic(P, desc->n_fft, di, st);
  }
  return (int)e;
}

}  // namespace aip

using namespace aip;

extern "C" {

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
  P.reflect = desc->center == 2;
  if (P.reflect && L <= P.pad) return AIP_ERR_UNSUPPORTED;
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
  P.var_prefetch = tunables().var_no_prefetch ? 0 : 1;
  variant_meta_kernel<<<(unsigned)((P.B + 255) / 256), 256, 0, st>>>(gap_samples, static_cast<int4*>(workspace), P.B, (int)G, P);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const size_t fill_smem = (size_t)4 * kFillChunk * sizeof(float);
  const bool aligned4 = ((reinterpret_cast<uintptr_t>(mag_out) | reinterpret_cast<uintptr_t>(clean_mag)) & 3) == 0;
  if (tunables().var_no_fill) {
    // experiment switch: transform kernel only (tools/variant_probe.py)
  } else if (aligned4 && !tunables().var_fill_scalar) {
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

}  // extern "C"
