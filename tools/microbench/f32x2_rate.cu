// Microbenchmark: issue rate of packed FP32x2 (FADD2 / FFMA2) vs scalar FADD / FFMA on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu && ./f32x2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) rate(float2* out, int iters, float seed) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  const float2 c = make_float2(1.0000001f, 0.9999999f), d = make_float2(1e-7f, -1e-7f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, c.x, d.x); a[i].y = fmaf(a[i].y, c.y, d.y); }        // 2 FFMA
      if (MODE == 1) { a[i] = __ffma2_rn(a[i], c, d); }                                           // 1 FFMA2
      if (MODE == 2) { a[i].x = a[i].x + d.x; a[i].y = a[i].y + d.y; }                            // 2 FADD
      if (MODE == 3) { a[i] = __fadd2_rn(a[i], d); }                                              // 1 FADD2
      if (MODE == 4) { a[i] = __ffma2_rn(a[i], c, d); a[i].x = a[i].x + d.y; }                    // FFMA2 + FADD
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int ops_per_iter_per_thread) {
  float2* out;
  const int blocks = 148 * 2, threads = 512, iters = 20000;
  cudaMalloc(&out, sizeof(float2) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate<MODE><<<blocks, threads>>>(out, 100, 1.0f);
  cudaEventRecord(e0);
  rate<MODE><<<blocks, threads>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop_lanes = (double)blocks * threads * iters * ops_per_iter_per_thread;   // fp32 lane-ops
  printf("%-18s %8.3f ms  %7.2f T lane-op/s  (%.1f lane-ops/clk/SM at 1.965 GHz)\n", name, ms, flop_lanes / ms * 1e-9,
         flop_lanes / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  run<0>("2x FFMA", 16);
  run<1>("1x FFMA2", 16);
  run<2>("2x FADD", 16);
  run<3>("1x FADD2", 16);
  run<4>("FFMA2 + FADD", 24);
  return 0;
}
