// Microbenchmark: issue rate of FP32x2 operand forms that matter for twiddle-fused butterflies on sm_100a:
// FADD2 reg,reg / FFMA2 reg,-1,reg (the exact subtraction) / FFMA2 reg,imm,reg / FFMA2 reg,reg,reg / FMUL2 reg,imm.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_imm f32x2_imm.cu && ./f32x2_imm
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) rate(float2* out, int iters, float seed) {
  float2 a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed - i); b[i] = out[(threadIdx.x * 8 + i) & 1023]; }
  const float2 w = make_float2(seed * 0.5f, seed * 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = __fadd2_rn(a[i], b[i]);
      if (MODE == 1) a[i] = __ffma2_rn(b[i], make_float2(-1.0f, -1.0f), a[i]);
      if (MODE == 2) a[i] = __ffma2_rn(b[i], make_float2(0.92387953f, 0.92387953f), a[i]);
      if (MODE == 3) a[i] = __ffma2_rn(b[i], w, a[i]);
      if (MODE == 4) a[i] = __fmul2_rn(a[i], make_float2(0.99999994f, 0.99999994f));
      if (MODE == 5) { a[i].x = fmaf(b[i].x, 0.92387953f, a[i].x); a[i].y = fmaf(b[i].y, 0.92387953f, a[i].y); }
      if (MODE == 6) { a[i].x = a[i].x + b[i].x; a[i].y = a[i].y + b[i].y; }
      if (MODE == 7) a[i] = __ffma2_rn(b[(i + 1) & 7], make_float2(0.92387953f, 0.92387953f), a[i]);
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chain, one warp per scheduler: time per instruction = its dependent-issue latency
template <int MODE>
__global__ void __launch_bounds__(128) lat(float2* out, int iters, float seed) {
  float2 a = make_float2(seed + threadIdx.x, seed), b = out[threadIdx.x & 1023];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a = __fadd2_rn(a, b);
      if (MODE == 1) a = __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
      if (MODE == 2) a = __ffma2_rn(a, make_float2(0.92387953f, 0.92387953f), b);
      if (MODE == 3) a = __fmul2_rn(a, make_float2(0.99999994f, 0.99999994f));
      if (MODE == 4) a.x = fmaf(a.x, 0.92387953f, b.x);
      if (MODE == 5) a.x = a.x + b.x;
      if (MODE == 6) a = __ffma2_rn(b, make_float2(0.92387953f, 0.92387953f), a);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
template <int MODE>
void runlat(const char* name) {
  float2* out;
  const int iters = 20000;
  cudaMalloc(&out, sizeof(float2) * 148 * 1024);
  cudaMemset(out, 0, sizeof(float2) * 148 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  lat<MODE><<<148, 128>>>(out, 100, 1.0f);
  cudaEventRecord(e0);
  lat<MODE><<<148, 128>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("latency %-24s %.3f ns per dependent instruction\n", name, ms * 1e6 / (16.0 * iters));
  cudaFree(out);
}

template <int MODE>
void run(const char* name) {
  float2* out;
  const int blocks = 148, threads = 512, iters = 20000;
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
  // warp-instructions per SM-sub-partition: 4 warps per scheduler x 8 x iters (x 2 for the scalar modes)
  const double winstr = 4.0 * 8 * iters * ((MODE == 5 || MODE == 6) ? 2 : 1);
  printf("%-28s %8.3f ms  %.3f ns per warp-instruction and scheduler\n", name, ms, ms * 1e6 / winstr);
  cudaFree(out);
}

int main() {
  run<0>("FADD2 r,r");
  run<1>("FFMA2 r,-1,r");
  run<2>("FFMA2 r,imm,r");
  run<3>("FFMA2 r,r,r");
  run<4>("FMUL2 r,imm");
  run<5>("2 x FFMA r,imm,r");
  run<6>("2 x FADD r,r");
  run<7>("FFMA2 r,imm,r (other bank)");
  runlat<0>("FADD2");
  runlat<1>("FFMA2 b,-1,a (acc dep)");
  runlat<2>("FFMA2 a,imm,b (mul dep)");
  runlat<3>("FMUL2 imm");
  runlat<4>("FFMA imm");
  runlat<5>("FADD");
  runlat<6>("FFMA2 b,imm,a (acc dep)");
  return 0;
}
