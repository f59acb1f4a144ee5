// Probe for the tensor-core formulation of the forward kernel's stage 2 (the 16-point DFTs over n1 of the 16 x 16 factorisation):
//   D[(k2, frame), (k1, c)] = sum_{(n1, c')} A[(k2, frame), (n1, c')] * B[(n1, c'), (k1, c)]
// as tcgen05.mma kind::tf32 with fp32 accuracy from a two-term split of both operands (A = Ahi + Alo, B = Bhi + Blo;
// Ahi Bhi + Alo Bhi + Ahi Blo).  Checks, on a B200:
//   * the shared-memory descriptor arithmetic (K-major, no swizzle, NON-canonical leading byte offset 144 B = padded core
//     matrices, which makes the producers' 16-byte stores conflict-free),
//   * the accumulator layout read back with tcgen05.ld.32x32b,
//   * the error against a float64 DFT, and the time per 32-frame tile of the MMA chain and of the TMEM read-back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_dft16_probe tc_dft16_probe.cu ; run: ./tc_dft16_probe [lbo]
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// quarter w of M-block b holds job k2 = kQ[w][b]: every quarter's set is closed under k2 -> (16 - k2) % 16, so the warp that
// owns lanes 32 w .. 32 w + 31 can read both members of every (k, 256 - k) pair
__constant__ int kQ[4][4] = {{0, 8, 1, 15}, {2, 14, 3, 13}, {4, 12, 5, 11}, {6, 10, 7, 9}};

constexpr int kChunks = 16;                 // K = 64 = 16 chunks of 4 floats: (re_hi, im_hi, re_lo, im_lo) of one n1
constexpr int kSboB = 16 * 128, kLboB = 128;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  return d;                                 // layout type 0 = no swizzle, base offset 0, lbo mode 0
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

#define TMEM_LD32(taddr, v) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
               : "r"(taddr))

// Y: [tiles][32 frames][16 k2][16 n1][2] stage-1 output (after the inter-stage twiddle); out: [tiles][32][16 k2][16 k1][2]
__global__ void __launch_bounds__(256, 1) probe(const float* __restrict__ Y, float* __restrict__ out, int tiles, int reps, int lbo,
                                                 long long* cycles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sbo = kChunks * lbo, blk = 16 * sbo;          // bytes: 8-row group stride, M-block stride
  unsigned char* A = sm;                                   // 4 M-blocks
  float* B1 = reinterpret_cast<float*>(sm + 4 * blk);      // (Bhi | Bhi): pass 1 = Ahi Bhi + Alo Bhi
  float* B2 = B1 + 32 * 64;                                // (Blo | 0):   pass 2 = Ahi Blo
  // ---- B: row n = (k1, c) [c = 0 re, 1 im], K index = 4 n1 + s, s = (re_hi, im_hi, re_lo, im_lo)
  for (int idx = tid; idx < 32 * 64; idx += blockDim.x) {
    const int n = idx >> 6, k = idx & 63, n1 = k >> 2, s = k & 3, k1 = n >> 1, c = n & 1;
    float sn, cs;
    sincospif(-(float)((n1 * k1) & 15) / 8.0f, &sn, &cs);         // W16^(n1 k1) = cs + j sn
    // out_re = in_re cs - in_im sn ; out_im = in_re sn + in_im cs
    const float coef = (c == 0) ? ((s & 1) == 0 ? cs : -sn) : ((s & 1) == 0 ? sn : cs);
    const float hi = __uint_as_float(__float_as_uint(coef) & 0xffffe000u), lo = coef - hi;
    const int off = (n & 7) * 4 + (n >> 3) * (kSboB / 4) + n1 * (kLboB / 4) + s;      // floats
    B1[off] = hi;
    B2[off] = (s < 2) ? lo : 0.0f;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base;
  // M = 128, N = 32, tf32 x tf32 -> f32, both K-major
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t phase = 0;
  long long t_mma = 0, t_ld = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // ---- A: row (w, frame) of block b holds job k2 = kQ[w][b]
    const float* y = Y + (long long)tile * 32 * 16 * 16 * 2;
    for (int idx = tid; idx < 32 * 16 * 16; idx += blockDim.x) {
      const int n1 = idx & 15, f = (idx >> 4) & 31, qi = idx >> 9;      // qi = 4 w + b
      const int w = qi >> 2, b = qi & 3, k2 = kQ[w][b];
      const float2 v = *reinterpret_cast<const float2*>(y + ((f * 16 + k2) * 16 + n1) * 2);
      const float hr = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), hi_ = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
      const int row = w * 32 + f;
      float4* dst = reinterpret_cast<float4*>(A + b * blk + (row & 7) * 16 + (row >> 3) * sbo + n1 * lbo);
      *dst = make_float4(hr, hi_, v.x - hr, v.y - hi_);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    long long c0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int b = 0; b < 4; ++b) {
          const uint32_t d = tbase + 32 * b;
          for (int pass = 0; pass < 2; ++pass) {
            const uint64_t a0 = make_desc(smem_u32(A + b * blk), lbo, sbo);
            const uint64_t b0 = make_desc(smem_u32(pass ? B2 : B1), kLboB, kSboB);
            for (int j = 0; j < 8; ++j)       // K = 8 tf32 = two 16-byte chunks per instruction
              mma_tf32(d, a0 + (uint64_t)((2 * j * lbo) >> 4), b0 + (uint64_t)((2 * j * kLboB) >> 4), idesc, (pass | j) ? 1u : 0u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long c1 = clock64();
    // ---- read back: warp (w = warp & 3, h = warp >> 2) takes blocks 2h and 2h + 1 of its lane quarter
    const int w = warp & 3, h = warp >> 2;
    uint32_t va[32], vb[32];
    for (int r = 0; r < reps; ++r) {
      TMEM_LD32(tbase + ((uint32_t)(32 * w) << 16) + 32 * (2 * h), va);
      TMEM_LD32(tbase + ((uint32_t)(32 * w) << 16) + 32 * (2 * h + 1), vb);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    long long c2 = clock64();
    t_mma += c1 - c0; t_ld += c2 - c1;
    float* o = out + ((long long)tile * 32 + lane) * 16 * 16 * 2;
    const int ka = kQ[w][2 * h], kb = kQ[w][2 * h + 1];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      o[ka * 32 + j] = __uint_as_float(va[j]);
      o[kb * 32 + j] = __uint_as_float(vb[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (tid == 0 && cycles) { cycles[2 * blockIdx.x] = t_mma; cycles[2 * blockIdx.x + 1] = t_ld; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}

// Plan 2: Ahi and Alo in separate operand tiles (K = 32 each); per M-block 4 MMAs Ahi x [Bhi | Blo] (N = 64) + 4 MMAs Alo x Bhi
// (N = 32): every A tile is read ONCE (the MMA is bound by its shared-memory operand reads: A = 4 KB per instruction).  The
// read-back adds columns 32..63 (Ahi Blo) to columns 0..31.
__global__ void __launch_bounds__(256, 1) probe2(const float* __restrict__ Y, float* __restrict__ out, int tiles, int reps, int lbo,
                                                  long long* cycles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sbo = 8 * lbo, blk = 16 * sbo;
  unsigned char* Ahi = sm;
  unsigned char* Alo = sm + 4 * blk;
  float* Bc = reinterpret_cast<float*>(sm + 8 * blk);      // 64 rows (Bhi rows 0..31, Blo rows 32..63) x K 32
  constexpr int lboB = 128, sboB = 8 * 128;
  for (int idx = tid; idx < 64 * 32; idx += blockDim.x) {
    const int n = idx >> 5, k = idx & 31, q = k >> 2, s = k & 3, n1 = 2 * q + (s >> 1), cin = s & 1;
    const int nn = n & 31, k1 = nn >> 1, c = nn & 1;
    float sn, cs;
    sincospif(-(float)((n1 * k1) & 15) / 8.0f, &sn, &cs);
    const float coef = (c == 0) ? (cin == 0 ? cs : -sn) : (cin == 0 ? sn : cs);
    const float hi = __uint_as_float(__float_as_uint(coef) & 0xffffe000u), lo = coef - hi;
    Bc[(n & 7) * 4 + (n >> 3) * (sboB / 4) + q * (lboB / 4) + s] = n < 32 ? hi : lo;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base;
  const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t phase = 0;
  long long t_mma = 0, t_ld = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const float* y = Y + (long long)tile * 32 * 16 * 16 * 2;
    for (int idx = tid; idx < 32 * 16 * 16; idx += blockDim.x) {
      const int n1 = idx & 15, f = (idx >> 4) & 31, qi = idx >> 9;
      const int w = qi >> 2, b = qi & 3, k2 = kQ[w][b];
      const float2 v = *reinterpret_cast<const float2*>(y + ((f * 16 + k2) * 16 + n1) * 2);
      const float hr = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), hi_ = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
      const int row = w * 32 + f;
      const int off = b * blk + (row & 7) * 16 + (row >> 3) * sbo + (n1 >> 1) * lbo + (n1 & 1) * 8;
      *reinterpret_cast<float2*>(Ahi + off) = make_float2(hr, hi_);
      *reinterpret_cast<float2*>(Alo + off) = make_float2(v.x - hr, v.y - hi_);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    long long c0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t b0 = make_desc(smem_u32(Bc), lboB, sboB);
        for (int b = 0; b < 4; ++b) {
          const uint32_t d = tbase + 64 * b;
          const uint64_t ah = make_desc(smem_u32(Ahi + b * blk), lbo, sbo), al = make_desc(smem_u32(Alo + b * blk), lbo, sbo);
          for (int j = 0; j < 4; ++j)
            mma_tf32(d, ah + (uint64_t)((2 * j * lbo) >> 4), b0 + (uint64_t)((2 * j * lboB) >> 4), idesc64, j ? 1u : 0u);
          for (int j = 0; j < 4; ++j)
            mma_tf32(d, al + (uint64_t)((2 * j * lbo) >> 4), b0 + (uint64_t)((2 * j * lboB) >> 4), idesc32, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long c1 = clock64();
    const int w = warp & 3, h = warp >> 2;
    uint32_t va[32], vb[32], xa[32], xb[32];
    for (int r = 0; r < reps; ++r) {
      TMEM_LD32(tbase + ((uint32_t)(32 * w) << 16) + 64 * (2 * h), va);
      TMEM_LD32(tbase + ((uint32_t)(32 * w) << 16) + 64 * (2 * h) + 32, xa);
      TMEM_LD32(tbase + ((uint32_t)(32 * w) << 16) + 64 * (2 * h + 1), vb);
      TMEM_LD32(tbase + ((uint32_t)(32 * w) << 16) + 64 * (2 * h + 1) + 32, xb);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    long long c2 = clock64();
    t_mma += c1 - c0; t_ld += c2 - c1;
    float* o = out + ((long long)tile * 32 + lane) * 16 * 16 * 2;
    const int ka = kQ[w][2 * h], kb = kQ[w][2 * h + 1];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      o[ka * 32 + j] = __uint_as_float(va[j]) + __uint_as_float(xa[j]);
      o[kb * 32 + j] = __uint_as_float(vb[j]) + __uint_as_float(xb[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (tid == 0 && cycles) { cycles[2 * blockIdx.x] = t_mma; cycles[2 * blockIdx.x + 1] = t_ld; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase) : "memory");
}

int main(int argc, char** argv) {
  const int lbo = argc > 1 ? atoi(argv[1]) : 144;
  const int plan = argc > 3 ? atoi(argv[3]) : 1;
  const int tiles = 148 * 2, reps = argc > 2 ? atoi(argv[2]) : 1;
  const size_t n = (size_t)tiles * 32 * 16 * 16 * 2;
  std::vector<float> hy(n);
  srand(1);
  for (auto& v : hy) v = (float)rand() / RAND_MAX - 0.5f;
  float *dy, *dout;
  long long* dcyc;
  CK(cudaMalloc(&dy, n * 4)); CK(cudaMalloc(&dout, n * 4)); CK(cudaMalloc(&dcyc, 148 * 2 * 8));
  CK(cudaMemcpy(dy, hy.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, n * 4));
  const int smem = plan == 1 ? 4 * 16 * 16 * lbo + 2 * 32 * 64 * 4 : 8 * 16 * 8 * lbo + 64 * 32 * 4;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  printf("plan %d, lbo %d B, smem %d B, reps %d\n", plan, lbo, smem, reps);
  if (plan == 1) probe<<<148, 256, smem>>>(dy, dout, tiles, reps, lbo, dcyc);
  else probe2<<<148, 256, smem>>>(dy, dout, tiles, reps, lbo, dcyc);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> ho(n);
  std::vector<long long> cyc(148 * 2);
  CK(cudaMemcpy(ho.data(), dout, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc.data(), dcyc, 148 * 2 * 8, cudaMemcpyDeviceToHost));
  double worst = 0, big = 0;
  for (int t = 0; t < tiles; t += 37)
    for (int f = 0; f < 32; ++f)
      for (int k2 = 0; k2 < 16; ++k2)
        for (int k1 = 0; k1 < 16; ++k1) {
          double re = 0, im = 0;
          for (int n1 = 0; n1 < 16; ++n1) {
            const float* p = &hy[((((size_t)t * 32 + f) * 16 + k2) * 16 + n1) * 2];
            const double a = -2.0 * M_PI * ((n1 * k1) & 15) / 16.0;
            re += p[0] * cos(a) - p[1] * sin(a);
            im += p[0] * sin(a) + p[1] * cos(a);
          }
          const float* q = &ho[((((size_t)t * 32 + f) * 16 + k2) * 16 + k1) * 2];
          worst = fmax(worst, fmax(fabs(q[0] - re), fabs(q[1] - im)));
          big = fmax(big, fmax(fabs(re), fabs(im)));
        }
  printf("max abs err %.3e, max |ref| %.3f, relative %.3e\n", worst, big, worst / big);
  printf("cycles per tile (CTA 0): mma chain + commit + wait %.0f, 2 x tcgen05.ld.x32 + wait %.0f (per rep)\n",
         (double)cyc[0] / 2 / reps, (double)cyc[1] / 2 / reps);
  return worst / big < 1e-5 ? 0 : 2;
}
