// Mel front / back-end: the filter-bank contraction of librosa.feature.melspectrogram (reference
// utils.py:268-277) and the pseudo-inverse projection of utils.mel_spectrogram_to_audio (utils.py:375-383).
//
// Layouts follow the rest of the library: spectrograms [B, F, T] and mel spectrograms [B, n_mels, T], T contiguous, so a warp
// works on 32 consecutive frames of one row and every global access is a coalesced 128-byte line.
//
//   mel_project_kernel   out[b, m, t] = sum_f basis[m, f] * S[b, f, t].  librosa's mel basis is triangular: row m is non-zero on
//                        one short bin range [f0, f1) (host-computed `bands`), and a bin belongs to at most two neighbouring
//                        rows, so the contraction reads the spectrogram about twice (the second time from L2) instead of
//                        n_mels times: HBM-bound streaming, not a GEMM.
//   mel_inverse_kernel   out[b, f, t] = sum_m inv_basis[f, m] * mel[b, m, t] (dense, K = n_mels <= 256), sqrt epilogue.
#include "aip_device.cuh"
#include "aip_host.h"

namespace aip {

constexpr int kMelRows = 8;       // mel rows per CTA (one per warp)
constexpr int kMelFrames = 128;   // frames per CTA: 4 per lane

// one warp = one mel row m, lanes along T (4 frames each, 32 apart: every load is a full line)
__global__ void __launch_bounds__(kMelRows * 32) mel_project_kernel(const float* __restrict__ basis, const int* __restrict__ bands,
                                                                   const float* __restrict__ spec, long long F, long long T,
                                                                   int n_mels, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.z;
  const int m = blockIdx.y * kMelRows + warp;
  if (m >= n_mels) return;
  const long long t0 = (long long)blockIdx.x * kMelFrames + lane;
  int f0 = bands[2 * m], f1 = bands[2 * m + 1];
  if (f0 < 0) f0 = 0;
  if (f1 > F) f1 = (int)F;
  const float* w = basis + (long long)m * F;
  const float* s = spec + b * F * T;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int f = f0; f < f1; ++f) {
    const float wf = __ldg(w + f);
    const float* row = s + (long long)f * T;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long t = t0 + 32 * j;
      if (t < T) acc[j] = fmaf(wf, row[t], acc[j]);
    }
  }
  float* o = out + (b * n_mels + m) * T;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long t = t0 + 32 * j;
    if (t < T) o[t] = acc[j];
  }
}

// CTA = 32 frames x 32 output bins; the [32, n_mels] slice of inv_basis sits in shared memory (pitch n_mels + 1), the mel
// column of a frame is read once per CTA row group: thread (ty, lane) accumulates bins ty, ty + 8, ty + 16, ty + 24.
constexpr int kInvBins = 32;
__global__ void __launch_bounds__(256) mel_inverse_kernel(const float* __restrict__ inv_basis, const float* __restrict__ mel,
                                                          long long F, long long T, int n_mels, int take_sqrt,
                                                          float* __restrict__ out) {
  extern __shared__ float wsm[];                 // [kInvBins][n_mels + 1]
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long b = blockIdx.z;
  const long long fbase = (long long)blockIdx.y * kInvBins;
  const long long t = (long long)blockIdx.x * 32 + lane;
  const int pitch = n_mels + 1;
  for (int idx = threadIdx.x; idx < kInvBins * n_mels; idx += blockDim.x) {
    const int r = idx / n_mels, m = idx - r * n_mels;
    wsm[r * pitch + m] = (fbase + r < F) ? inv_basis[(fbase + r) * n_mels + m] : 0.0f;
  }
  __syncthreads();
  if (t >= T) return;
  const float* col = mel + b * n_mels * T + t;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int m = 0; m < n_mels; ++m) {
    const float v = col[(long long)m * T];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fmaf(wsm[(ty + 8 * j) * pitch + m], v, acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long f = fbase + ty + 8 * j;
    if (f < F) out[(b * F + f) * T + t] = acc[j];
  }
}

// np.sqrt of the projection (utils.py:381-383), in place.  A clearly negative projection becomes NaN exactly as in the reference.
// A negative value BELOW THE NOISE FLOOR of an fp32 power spectrogram -- |v| <= kNegGuard x the clip's largest projected power;
// an fp32 FFT leaves ~1e-7 of the peak amplitude in every bin, i.e. ~1e-14 of the peak power, and the pseudo-inverse amplifies
// it by up to a few hundred -- is rounding residue of a projection whose exact value is >= 0 (librosa's float64-internal STFT
// has a smooth leakage floor there and the reference gets a small positive number): it is taken as zero instead of
// poisoning the whole waveform with NaN.  Documented deviation.
constexpr float kNegGuard = 1e-9f;
__global__ void sqrt_guard_kernel(float* x, long long n_per_clip, long long B, const float* peaks) {
  const long long total = B * n_per_clip;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float floor_v = -kNegGuard * peaks[i / n_per_clip];
    x[i] = (v < 0.0f && v >= floor_v) ? 0.0f : sqrtf(v);
  }
}

}  // namespace aip

using namespace aip;

extern "C" {

int aip_mel_project_f32(const float* basis, const int32_t* bands, const float* spec_pow, int64_t B, int64_t F, int64_t T,
                        int64_t n_mels, float* mel_out, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!basis || !bands || !spec_pow || !mel_out || B < 0 || F < 1 || T < 0 || n_mels < 1) return AIP_ERR_ARG;
  if (B > 65535 || n_mels > 65535 * kMelRows || F > 0x7fffffffLL) return AIP_ERR_UNSUPPORTED;
  if (B == 0 || T == 0) return AIP_OK;
  const dim3 grid((unsigned)((T + kMelFrames - 1) / kMelFrames), (unsigned)((n_mels + kMelRows - 1) / kMelRows), (unsigned)B);
  mel_project_kernel<<<grid, kMelRows * 32, 0, static_cast<cudaStream_t>(stream)>>>(basis, bands, spec_pow, F, T, (int)n_mels, mel_out);
  return (int)cudaGetLastError();
}

int aip_mel_inverse_f32(const float* inv_basis, const float* mel, int64_t B, int64_t F, int64_t T, int64_t n_mels,
                        int32_t take_sqrt, float* out, float* peaks, void* stream) {
  const DevInfo di = dev_info();
  if (!di.ok) return AIP_ERR_DEVICE;
  if (!inv_basis || !mel || !out || B < 0 || F < 1 || T < 0 || n_mels < 1 || (take_sqrt && !peaks)) return AIP_ERR_ARG;
  if (B > 65535 || n_mels > 1024 || (F + kInvBins - 1) / kInvBins > 65535) return AIP_ERR_UNSUPPORTED;
  if (B == 0 || T == 0) return AIP_OK;
  const size_t smem = (size_t)kInvBins * (size_t)(n_mels + 1) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mel_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid((unsigned)((T + 31) / 32), (unsigned)((F + kInvBins - 1) / kInvBins), (unsigned)B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mel_inverse_kernel<<<grid, 256, smem, st>>>(inv_basis, mel, F, T, (int)n_mels, 0, out);
  e = cudaGetLastError();
  if (e != cudaSuccess || !take_sqrt) return (int)e;
  e = cudaMemsetAsync(peaks, 0, (size_t)B * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  e = launch_peak(out, F * T, B, F * T, peaks, st);
  if (e != cudaSuccess) return (int)e;
  sqrt_guard_kernel<<<ew_grid(B * F * T, di.sms), 256, 0, st>>>(out, F * T, B, peaks);
  return (int)cudaGetLastError();
}

}  // extern "C"
