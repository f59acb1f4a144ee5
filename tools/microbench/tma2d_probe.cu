// Probe: the canonical 2-D TMA tile load (CUDA programming guide layout), int32 data.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int W = 64, H = 16;
template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap map, int* out, int c0, int c1) {
  __shared__ alignas(128) int buf[H * W];
  __shared__ alignas(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(H * W * 4) : "memory");
    if (MODE == 0)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(buf)), "l"(&map), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(buf)), "l"(&map), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int GW = 1024, GH = 512;
  std::vector<int> h((size_t)GW * GH);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (int)i;
  int *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, H * W * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t ge = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fp, 12000, cudaEnableDefault, &q);
  printf("entry %p err %d status %d\n", fp, (int)ge, (int)q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap map;
  const cuuint64_t dims[2] = {GW, GH};
  const cuuint64_t strides[1] = {GW * 4ull};
  const cuuint32_t box[2] = {W, H};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("2d encode -> %d; map bytes:", (int)r);
  for (int i = 0; i < 16; ++i) printf(" %016llx", ((unsigned long long*)&map)[i]);
  printf("\n");
  if (mode == 0) probe<0><<<1, 128>>>(map, o, 128, 32); else probe<1><<<1, 128>>>(map, o, 128, 32);
  cudaError_t e = cudaDeviceSynchronize();
  printf("2d mode %d kernel -> %s\n", mode, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<int> got(H * W);
  cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) if (got[y * W + x] != h[(size_t)(32 + y) * GW + 128 + x]) ++bad;
  printf("2d mismatches %d\n", bad);
  return 0;
}
