// Microbenchmark: cost of the operand FORMS of packed FP32x2 math on sm_100a (plain, half-swapped, negated,
// three distinct register pairs, squared operand) against scalar math, at the occupancy of the STFT kernels
// (16 warps / SM, one 512-thread CTA per SM) and as a dependent chain (latency).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_forms f32x2_forms.cu && ./f32x2_forms
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 swp(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 ng(float2 a) { return make_float2(-a.x, -a.y); }

// MODE: see names[] below.  ILP independent accumulators a[i]; b[i], c[i] are distinct register pairs.
template <int MODE, int ILP>
__global__ void __launch_bounds__(512, 1) forms(float2* out, int iters, float seed) {
  float2 a[ILP], b[ILP], c[ILP];
  // run-time operand values (read back from `out`, which the host zero-fills): every a/b/c is a live register pair
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    const float2 z = out[(threadIdx.x + 37 * i) & 511];
    a[i] = make_float2(seed + i + threadIdx.x + z.x, seed - i + z.y);
    b[i] = make_float2(1.0f + z.x * (i + 1), 1.0f - z.y * (i + 2));
    c[i] = make_float2(z.x * (i + 3), z.y - z.x * (i + 5));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const int j = (i + 1) % ILP;
      if (MODE == 0) { a[i].x = fmaf(a[i].x, b[i].x, c[i].x); a[i].y = fmaf(a[i].y, b[i].y, c[i].y); }   // 2 FFMA, 3 regs
      if (MODE == 1) a[i] = __ffma2_rn(a[i], b[i], c[i]);                                                 // FFMA2, 3 pairs
      if (MODE == 2) a[i] = __fadd2_rn(a[i], c[i]);                                                       // FADD2 plain
      if (MODE == 3) a[i] = __fadd2_rn(a[i], swp(c[i]));                                                  // FADD2 LO_HI
      if (MODE == 4) a[i] = __fadd2_rn(a[i], ng(c[i]));                                                   // FADD2 negated
      if (MODE == 5) a[i] = __ffma2_rn(c[i], make_float2(-1.0f, -1.0f), a[i]);                            // sub via FFMA2
      if (MODE == 6) { a[i].x = a[i].x + c[i].x; a[i].y = a[i].y + c[i].y; }                              // 2 FADD
      if (MODE == 7) a[i] = __ffma2_rn(a[i], make_float2(b[i].x, b[i].x), c[i]);                          // FFMA2 broadcast b
      if (MODE == 8) a[i] = __ffma2_rn(ng(a[i]), b[i], c[i]);                                             // FFMA2 negated a
      if (MODE == 9) a[i] = __ffma2_rn(a[i], a[i], c[i]);                                                 // FFMA2 squared
      if (MODE == 10) a[i] = __ffma2_rn(a[j], b[i], a[i]);                                                // FFMA2 cross (fft-like)
      if (MODE == 11) { a[i].x = fmaf(a[j].x, b[i].x, a[i].x); a[i].y = fmaf(a[j].y, b[i].y, a[i].y); }   // 2 FFMA cross
      if (MODE == 12) { a[i].x = __fsqrt_rn(a[i].x) ; asm volatile("" ::: "memory"); }                    // placeholder
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MUFU throughput / mix: 1 MUFU + N FFMA2 per step
template <int NF>
__global__ void __launch_bounds__(512, 1) mufu_mix(float2* out, int iters, float seed) {
  float2 a[8];
  float m[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = seed + i + threadIdx.x;
  const float2 c = make_float2(1.0000001f, 0.9999999f), d = make_float2(1e-7f, -1e-7f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
#pragma unroll
      for (int q = 0; q < NF; ++q) a[(i * NF + q) & 7] = __ffma2_rn(a[(i * NF + q) & 7], c, d);
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
#pragma unroll
  for (int i = 0; i < 4; ++i) s.x += m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float time_kernel(K launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(50);
  cudaEventRecord(e0);
  launch(4000);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

static float2* g_out;
static const char* names[] = {"2 FFMA (3 regs)", "FFMA2 (3 pairs)", "FADD2 plain", "FADD2 LO_HI", "FADD2 negated",
                              "sub as FFMA2(-1)", "2 FADD", "FFMA2 bcast b", "FFMA2 negated a", "FFMA2 squared",
                              "FFMA2 cross", "2 FFMA cross"};

template <int MODE, int ILP>
void run(int threads, int smem, const char* tag) {
  auto k = forms<MODE, ILP>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const float ms = time_kernel([&](int it) { cudaMemsetAsync(g_out, 0, sizeof(float2) * 148 * 512); k<<<148, threads, smem>>>(g_out, it, 1.0f); });
  const double steps = 148.0 * threads * 4000.0 * ILP;            // "steps" = one MODE statement (2 fp32 lanes of work)
  const double clk = ms * 1e-3 * 1.965e9;
  printf("%-18s ILP %d %4d thr %-8s %8.3f ms  %6.1f lane-ops/clk/SM   %5.2f clk per step per warp-slot\n", names[MODE], ILP,
         threads, tag, ms, 2.0 * steps / clk / 148.0, clk / (4000.0 * ILP));
}

template <int MODE>
void all() {
  run<MODE, 8>(512, 200 * 1024, "thrput");
  run<MODE, 2>(512, 200 * 1024, "ilp2");
  run<MODE, 1>(32, 0, "latency");
}

template <int NF>
void run_mix() {
  auto k = mufu_mix<NF>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const float ms = time_kernel([&](int it) { k<<<148, 512, 200 * 1024>>>(g_out, it, 1.0f); });
  const double clk = ms * 1e-3 * 1.965e9;
  printf("MUFU.LG2 + %d FFMA2: %8.3f ms  %6.2f clk per (MUFU + %d FFMA2) per SMSP-warp  -> %5.2f MUFU/clk/SM\n", NF, ms,
         clk / (4000.0 * 4) / 4.0, NF, 148.0 * 512 * 4000.0 * 4 / clk / 148.0);
}

int main() {
  cudaMalloc(&g_out, sizeof(float2) * 148 * 512);
  cudaMemset(g_out, 0, sizeof(float2) * 148 * 512);
  all<0>(); all<1>(); all<2>(); all<3>(); all<4>(); all<5>(); all<6>(); all<7>(); all<8>(); all<9>(); all<10>(); all<11>();
  run_mix<0>(); run_mix<1>(); run_mix<2>(); run_mix<4>(); run_mix<8>();
  return 0;
}
