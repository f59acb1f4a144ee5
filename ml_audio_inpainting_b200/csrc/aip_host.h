// Host-side helpers shared by the translation units behind the C ABI (include/aip_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/aip_b200.h"
#include "aip_tiles.cuh"

namespace aip {

struct DevInfo { int ok; int sms; int max_smem; int max_smem_sm; };

static DevInfo dev_info() {
  DevInfo d{0, 0, 0, 0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return d;
  int major = 0, minor = -1;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&d.max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  // the library holds sm_100a SASS only, and arch-specific ("a") code runs on exactly that compute capability: 10.0.
  // Any other 10.x part (sm_103, ...) would fail every launch with "no kernel image"; it gets AIP_ERR_DEVICE instead.
  d.ok = (major == 10 && minor == 0);
  return d;
}

static inline int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
static inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
static inline int ew_grid(long long total, int sms) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)sms * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static long long num_frames(long long L, int n_fft, int hop, int center) {
  if (L < 0 || n_fft <= 0 || hop <= 0) return -1;
  const long long padded = L + (center ? 2LL * (n_fft / 2) : 0);
  if (padded < n_fft) return -1;
  return 1 + (padded - n_fft) / hop;
}

static long long istft_length(long long T, int n_fft, int hop, int center, long long length) {
  if (length > 0) return length;
  long long n = (long long)n_fft + (long long)hop * (T - 1);
  if (center) n -= 2LL * (n_fft / 2);
  return n;
}

static long long istft_used_frames(long long T, int n_fft, int hop, int center, long long length) {
  if (length <= 0) return T;
  const long long padded = length + (center ? 2LL * (n_fft / 2) : 0);
  const long long nf = (padded + hop - 1) / hop;
  return nf < T ? nf : T;
}

// Experiment switches (A/B timing, tests of alternative code paths).  The AIP_* environment variables are read ONCE, when the
// library is loaded (and again only on an explicit aip_debug_reload_env()): no getenv() on any launch path.
struct Tunables {
  int fwd_tile_bufs;      // AIP_FWD_TILE_BUFS   1..3: cap on the ring of staged-waveform buffers (0 = no cap)
  int fwd_no_shape;       // AIP_FWD_NO_SHAPE    1: never pick the shape-specialised (T_out 417 / 834) forward builds
  int fwd_chunk;          // AIP_FWD_CHUNK       tiles per draw of the dynamic schedule (0 = built-in choice)
  int inv_tma;            // AIP_INV_TMA         0: never stage the inverse kernel's rows with TMA tensor boxes (default: where legal)
  int ola_fast_mask;      // AIP_OLA_FAST        bit mask of the specialised overlap-adds that may be used (-1 = all)
  int inv_l2_prefetch;    // AIP_INV_L2_PREFETCH 0: stage A of the inverse kernels does not request the next tile's rows from L2
  int inv_bufs;           // AIP_INV_BUFS        exchange buffers in the inverse ring (0 = built-in choice)
  int gl_unfused;         // AIP_GL_UNFUSED      1: Griffin-Lim with the separate phase-update kernel
  int var_no_prefetch;    // AIP_VAR_NO_PREFETCH 1: gap-variant tiles do not request their rows from L2 ahead of the stores
  int var_fill_scalar;    // AIP_VAR_FILL=scalar gap-variant copy pass with store instructions instead of bulk copies
  int var_no_fill;        // AIP_VAR_NO_FILL     1: skip the copy pass (timing the transform kernel alone)
  int pow2_ola_fast;      // AIP_POW2_OLA_FAST   0: the tiled inverse always uses the general overlap-add gather
  int pow2_span;          // AIP_POW2_SPAN       0: the tiled forward kernel loads every frame from global instead of staging the tile's span
  int pow2;               // AIP_POW2            0: n_fft != 512 runs the one-frame-per-CTA radix-2 kernels instead of the tiled radix-16 ones
};
const Tunables& tunables();
void tunables_reload();

// ---- cross-unit entry points (defined in aip_fwd.cu / aip_pow2.cu / aip_misc.cu) ---------------------------------------------------
int run_fwd(const aip_stft_desc* desc, FwdParams P, long long T_out, cudaStream_t st);
bool fwd_fast_ok(const aip_stft_desc* d, const DevInfo& di);
bool pow2_ok(int n_fft);
cudaError_t launch_fwd_pow2(FwdParams P, int n_fft, const DevInfo& di, cudaStream_t st);
bool pow2_ola_ok(int n_fft, int hop);
cudaError_t launch_inv_pow2(InvParams P, int n_fft, float* frames, const DevInfo& di, cudaStream_t st);
cudaError_t launch_peak(const float* in, long long pitch, long long B, long long L, float* peaks, cudaStream_t st);
cudaError_t launch_pcm16(const float* in, long long in_pitch, short* pcm, long long pcm_pitch, long long B, long long L,
                         const float* peaks, int sms, cudaStream_t st);
cudaError_t launch_peak_scale(const float* in, long long in_pitch, float* out, long long out_pitch, long long B, long long L,
                              const float* peaks, int sms, cudaStream_t st);

}  // namespace aip
