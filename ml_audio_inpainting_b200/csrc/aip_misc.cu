// Small elementwise / reduction kernels (gap, mask, dB heuristic, peak normalisation) and the remaining C ABI entry points.
#include "aip_device.cuh"
#include "aip_host.h"

namespace aip {

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

// Row-wise elementwise passes: blockIdx.y walks the clips, blockIdx.x the samples -- no 64-bit division per element (the first
// versions split a flat index with i / L and i % L: 0.24 ms for a pass whose traffic takes 0.1 ms).
constexpr int kRowChunk = 256 * 8;          // samples per CTA and trip
static dim3 row_grid(long long B, long long L, int per_thread) {
  const long long chunk = (long long)kRowChunk * per_thread;
  long long gx = (L + chunk - 1) / chunk;
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  return dim3((unsigned)gx, (unsigned)(B < 65535 ? B : 65535));
}

__global__ void __launch_bounds__(256) gap_zero_kernel(const float* in, long long in_pitch, float* out, long long out_pitch,
                                                       long long B, long long L, const int* gaps) {
  for (long long b = blockIdx.y; b < B; b += gridDim.y) {
    const int g0 = gaps ? gaps[2 * b] : 0, g1 = gaps ? gaps[2 * b + 1] : 0;
    const float* x = in + b * in_pitch;
    float* y = out + b * out_pitch;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < L; s += (long long)gridDim.x * blockDim.x)
      y[s] = (s >= g0 && s < g1) ? 0.0f : x[s];
  }
}

__global__ void __launch_bounds__(256) gap_mask_kernel(float* mask, long long pitch, long long B, long long L, const int* gaps) {
  for (long long b = blockIdx.y; b < B; b += gridDim.y) {
    const int g0 = gaps ? gaps[2 * b] : 0, g1 = gaps ? gaps[2 * b + 1] : 0;
    float* y = mask + b * pitch;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < L; s += (long long)gridDim.x * blockDim.x)
      y[s] = (s >= g0 && s < g1) ? 0.0f : 1.0f;
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

__global__ void __launch_bounds__(256) peak_scale_kernel(const float* in, long long in_pitch, float* out, long long out_pitch,
                                                         long long B, long long L, const float* peaks) {
  for (long long b = blockIdx.y; b < B; b += gridDim.y) {
    const float pk = peaks[b];
    const float* x = in + b * in_pitch;
    float* y = out + b * out_pitch;
    if (pk < kFltMin) {                    // librosa.util.normalize leaves a clip whose peak is below tiny unchanged
      if (x != y)
        for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < L; s += (long long)gridDim.x * blockDim.x) y[s] = x[s];
      continue;
    }
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < L; s += (long long)gridDim.x * blockDim.x)
      y[s] = __fdiv_rn(x[s], pk);
  }
}

// save_audio's tail for the 16-bit FLAC it writes (utils.py:83-87): x / peak (librosa.util.normalize; peaks == nullptr: as is), then
// libsndfile's float -> PCM_16 conversion for FLAC: x * 32768, round half to even, clip to [-32768, 32767].  One thread = two samples.
__global__ void __launch_bounds__(256) pcm16_kernel(const float* in, long long in_pitch, short* pcm, long long pcm_pitch, long long B,
                                                    long long L, const float* peaks) {
  const bool vec = ((in_pitch | pcm_pitch) & 1) == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(pcm) & 3) == 0;
  for (long long b = blockIdx.y; b < B; b += gridDim.y) {
    const float pk = peaks ? peaks[b] : 0.0f;
    const float* x = in + b * in_pitch;
    short* q = pcm + b * pcm_pitch;
    for (long long s = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); s < L; s += 2LL * gridDim.x * blockDim.x) {
      const bool two = s + 1 < L;
      float v0, v1 = 0.0f;
      if (vec && two) {
        const float2 v = *reinterpret_cast<const float2*>(x + s);
        v0 = v.x; v1 = v.y;
      } else {
        v0 = x[s];
        if (two) v1 = x[s + 1];
      }
      if (pk >= kFltMin) { v0 = __fdiv_rn(v0, pk); v1 = __fdiv_rn(v1, pk); }
      const int q0 = min(32767, max(-32768, __float2int_rn(v0 * 32768.0f)));
      const int q1 = min(32767, max(-32768, __float2int_rn(v1 * 32768.0f)));
      if (vec && two) {
        *reinterpret_cast<short2*>(q + s) = make_short2((short)q0, (short)q1);
      } else {
        q[s] = (short)q0;
        if (two) q[s + 1] = (short)q1;
      }
    }
  }
}

cudaError_t launch_pcm16(const float* in, long long in_pitch, short* pcm, long long pcm_pitch, long long B, long long L,
                         const float* peaks, int sms, cudaStream_t st) {
  (void)sms;
  pcm16_kernel<<<row_grid(B, L, 2), 256, 0, st>>>(in, in_pitch, pcm, pcm_pitch, B, L, peaks);
  return cudaGetLastError();
}

cudaError_t launch_peak(const float* in, long long pitch, long long B, long long L, float* peaks, cudaStream_t st) {
  long long gx = (L + 256 * 8 - 1) / (256 * 8);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  peak_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, st>>>(in, pitch, L, peaks);
  return cudaGetLastError();
}

cudaError_t launch_peak_scale(const float* in, long long in_pitch, float* out, long long out_pitch, long long B, long long L,
                              const float* peaks, int sms, cudaStream_t st) {
  (void)sms;
  peak_scale_kernel<<<row_grid(B, L, 1), 256, 0, st>>>(in, in_pitch, out, out_pitch, B, L, peaks);
  return cudaGetLastError();
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}
static Tunables read_tunables() {
  Tunables t{};
  t.fwd_tile_bufs = env_int("AIP_FWD_TILE_BUFS", 0);
  t.fwd_no_shape = getenv("AIP_FWD_NO_SHAPE") ? 1 : 0;
  t.fwd_chunk = env_int("AIP_FWD_CHUNK", 0);
  t.inv_tma = env_int("AIP_INV_TMA", 1);
  t.ola_fast_mask = env_int("AIP_OLA_FAST", -1);
  t.inv_bufs = env_int("AIP_INV_BUFS", 0);
  t.inv_l2_prefetch = env_int("AIP_INV_L2_PREFETCH", 1);
  t.gl_unfused = getenv("AIP_GL_UNFUSED") ? 1 : 0;
  t.var_no_prefetch = getenv("AIP_VAR_NO_PREFETCH") ? 1 : 0;
  const char* f = getenv("AIP_VAR_FILL");
  t.var_fill_scalar = (f && f[0] == 's') ? 1 : 0;
  t.var_no_fill = getenv("AIP_VAR_NO_FILL") ? 1 : 0;
  t.pow2 = env_int("AIP_POW2", 1);
  t.pow2_span = env_int("AIP_POW2_SPAN", 1);
  t.pow2_ola_fast = env_int("AIP_POW2_OLA_FAST", 1);
  return t;
}
static Tunables g_tunables = read_tunables();        // once, when the shared object is loaded
const Tunables& tunables() { return g_tunables; }
void tunables_reload() { g_tunables = read_tunables(); }

}  // namespace aip

using namespace aip;

extern "C" {

int64_t aip_num_frames(int64_t L, int32_t n_fft, int32_t hop, int32_t center) {
  return num_frames(L, n_fft, hop, center);
}

int64_t aip_istft_length(int64_t T, int32_t n_fft, int32_t hop, int32_t center, int64_t length) {
  if (T < 1 || n_fft <= 0 || hop <= 0 || length < 0) return -1;
  return istft_length(T, n_fft, hop, center, length);
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
  gap_zero_kernel<<<row_grid(B, L, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, in_pitch, out, out_pitch, B, L, gap_samples);
  return (int)cudaGetLastError();
}

int aip_gap_mask_f32(float* mask, int64_t pitch, int64_t B, int64_t L, const int32_t* gap_samples, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!mask || B < 0 || L < 0 || pitch < L) return AIP_ERR_ARG;
  if (B * L == 0) return AIP_OK;
  gap_mask_kernel<<<row_grid(B, L, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, pitch, B, L, gap_samples);
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
  return (int)launch_peak_scale(in, in_pitch, out, out_pitch, B, L, peaks, di.sms, st);
}

int aip_wave_to_pcm16_f32(const float* in, int64_t in_pitch, int16_t* pcm, int64_t pcm_pitch, int64_t B, int64_t L,
                          int32_t peaks_mode, float* peaks, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!in || !pcm || B < 0 || L < 0 || in_pitch < L || pcm_pitch < L) return AIP_ERR_ARG;
  if (peaks_mode < AIP_PCM_RAW || peaks_mode > AIP_PCM_PEAKS_GIVEN || (peaks_mode != AIP_PCM_RAW && !peaks)) return AIP_ERR_ARG;
  if (B * L == 0) return AIP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (peaks_mode == AIP_PCM_NORMALIZE) {
    cudaError_t e = cudaMemsetAsync(peaks, 0, (size_t)B * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    for (long long lo = 0; lo < B; lo += 65535) {          // grid.y limit
      const long long n = B - lo < 65535 ? B - lo : 65535;
      e = launch_peak(in + lo * in_pitch, in_pitch, n, L, peaks + lo, st);
      if (e != cudaSuccess) return (int)e;
    }
  }
  return (int)launch_pcm16(in, in_pitch, reinterpret_cast<short*>(pcm), pcm_pitch, B, L,
                           peaks_mode == AIP_PCM_RAW ? nullptr : peaks, di.sms, st);
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

void aip_debug_reload_env(void) { tunables_reload(); }

const char* aip_version(void) { return "aip_b200 0.1.0 sm_100a"; }

int aip_device_supported(void) { return dev_info().ok; }

}  // extern "C"
