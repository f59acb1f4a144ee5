// Probe: plain 3-D TMA tile load (2t, k, b) of a complex [B, 257, T] spectrogram -- does the basic path work here?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int c0, int c1, int c2, int rows) {
  extern __shared__ __align__(128) float buf[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(rows * 256) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(buf)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int B = 3, T = 834, F = 257, rows = 16;
  std::vector<float> h((size_t)B * F * T * 2);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, rows * 64 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap map;
  // variant 0: rows = consecutive bins (stride 8T); variant 1: rows = every 16th bin via dim1 stride 128 T
  const cuuint64_t dims[3] = {2ull * T, variant == 0 ? (cuuint64_t)F : 16ull, (cuuint64_t)B};
  const cuuint64_t strides[2] = {variant == 0 ? 8ull * T : 128ull * T, 8ull * F * T};
  const cuuint32_t box[3] = {64u, (cuuint32_t)rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("3d variant %d encode -> %d\n", variant, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  const int t0 = argc > 2 ? atoi(argv[2]) : 57, k0 = 3, b = 1;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 256);
  probe<<<1, 128, rows * 256>>>(map, o, 2 * t0, variant == 0 ? k0 : 0, b, rows);
  cudaError_t e = cudaDeviceSynchronize();
  printf("3d variant %d kernel -> %s\n", variant, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> got(rows * 64);
  cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < rows; ++j)
    for (int x = 0; x < 64; ++x) {
      const int k = variant == 0 ? k0 + j : 16 * j;
      const size_t src = (((size_t)b * F + k) * T + t0) * 2 + x;
      if (got[j * 64 + x] != h[src]) ++bad;
    }
  printf("3d variant %d mismatches %d\n", variant, bad);
  return 0;
}
