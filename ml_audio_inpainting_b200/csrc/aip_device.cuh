// Device-side plumbing shared by the kernel translation units (aip_fwd.cu, aip_inv.cu, aip_mel.cu): mbarrier / TMA bulk-copy /
// named-barrier PTX wrappers and the shared-memory radix-2 FFT of the generic (n_fft != 512) kernels.
#pragma once
#include <cuda.h>            // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <stdint.h>

#include "aip_tiles.cuh"

namespace aip {

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
// defined time per attempt; a longer suspend-time hint measured SLOWER: later wake-ups; 1 - 2 dependent shared-memory
// loads between attempts, to poll less often, measured no different)
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

constexpr int kFwdThreads = 2 * kThreads;   // CTA of the n_fft = 512 kernels: 8 consumer + 8 producer warps
constexpr int kFwdTileBufs = 3;             // deepest ring of staged-waveform buffers (forward kernel)

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

}  // namespace aip
