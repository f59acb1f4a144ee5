// A host in C++ driving the C ABI directly -- no Python, no PyTorch: cudaMalloc'ed buffers, one stream, the round trip
// waveform -> STFT (log-magnitude of the gapped clip + complex spectrogram of the clean one) -> iSTFT, with the checks a binding's
// smoke test would make.  This is what a maintainer of a compiled service holding device buffers would write against
// include/aip_b200.h.
//
//   nvcc -std=c++17 -Iinclude -o abi_host_demo examples/abi_host_demo.cu -Lml_audio_inpainting_b200/lib -laip_b200 \
//        -Xlinker -rpath -Xlinker $PWD/ml_audio_inpainting_b200/lib
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "aip_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define AIP(x) do { int s_ = (x); if (s_ != AIP_OK) { std::printf("%s -> %s\n", #x, aip_status_string(s_)); return 3; } } while (0)

int main() {
  if (!aip_device_supported()) { std::printf("no sm_100 device: %s\n", aip_status_string(AIP_ERR_DEVICE)); return 1; }
  const int n_fft = 512, hop = 192, win_length = 384, B = 64;
  const long long L = 80000;                                      // 5 s at 16 kHz (the reference's max_len)
  const double kPi = 3.14159265358979323846;
  // librosa.filters.get_window("hann", 384, fftbins=True) centre-padded to n_fft (librosa.util.pad_center)
  std::vector<float> window(n_fft, 0.0f);
  for (int i = 0; i < win_length; ++i) window[(n_fft - win_length) / 2 + i] = (float)(0.5 - 0.5 * std::cos(2.0 * kPi * i / win_length));
  std::vector<float> wave((size_t)B * L);
  unsigned s = 12345u;
  for (auto& v : wave) { s = s * 1664525u + 1013904223u; v = 0.2f * ((float)(s >> 8) / 8388608.0f - 1.0f); }
  std::vector<int> gaps(2 * B);
  for (int b = 0; b < B; ++b) { gaps[2 * b] = 20000 + 100 * b; gaps[2 * b + 1] = gaps[2 * b] + 3200; }      // 0.2 s

  const long long T = aip_num_frames(L, n_fft, hop, 1), F = n_fft / 2 + 1, n_out = aip_istft_length(T, n_fft, hop, 1, 0);
  float *d_win, *d_wave, *d_mag, *d_spec, *d_wss, *d_out;
  int* d_gaps;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CK(cudaMalloc(&d_win, n_fft * sizeof(float)));
  CK(cudaMalloc(&d_wave, wave.size() * sizeof(float)));
  CK(cudaMalloc(&d_gaps, gaps.size() * sizeof(int)));
  CK(cudaMalloc(&d_mag, (size_t)B * F * T * sizeof(float)));
  CK(cudaMalloc(&d_spec, (size_t)B * F * T * 2 * sizeof(float)));
  CK(cudaMalloc(&d_wss, n_out * sizeof(float)));
  CK(cudaMalloc(&d_out, (size_t)B * n_out * sizeof(float)));
  CK(cudaMemcpyAsync(d_win, window.data(), n_fft * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_wave, wave.data(), wave.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_gaps, gaps.data(), gaps.size() * sizeof(int), cudaMemcpyHostToDevice, st));

  aip_stft_desc desc = {n_fft, hop, 1, win_length, d_win};
  // models/CNNBLSTM/dataset.py:103-106: log10(|STFT(gapped audio)| + 1e-9)
  AIP(aip_stft_fwd_f32(&desc, d_wave, B, L, L, d_gaps, nullptr, nullptr, 1, AIP_MAG_LOG10_EPS, 1e-9f, 1.0f, T, nullptr, d_mag, nullptr,
                       nullptr, st));
  // utils.extract_spectrogram of the clean audio, then utils.spectrogram_to_audio(S, phase_info=True)
  AIP(aip_stft_fwd_f32(&desc, d_wave, B, L, L, nullptr, nullptr, nullptr, 1, AIP_MAG_NONE, 0.0f, 1.0f, T, d_spec, nullptr, nullptr,
                       nullptr, st));
  AIP(aip_inv_window_sumsquare_f32(&desc, T, 0, d_wss, n_out, st));
  AIP(aip_istft_f32(&desc, d_spec, nullptr, nullptr, AIP_DOM_LINEAR, nullptr, B, T, 0, d_wss, d_out, n_out, nullptr, 0, st));
  std::vector<float> out((size_t)B * n_out), mag0((size_t)F * T);
  CK(cudaMemcpyAsync(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(mag0.data(), d_mag, mag0.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));

  double num = 0.0, den = 0.0;
  for (int b = 0; b < B; ++b)
    for (long long i = 0; i < n_out; ++i) {
      const double x = wave[(size_t)b * L + i], d = out[(size_t)b * n_out + i] - x;
      num += x * x; den += d * d;
    }
  const double snr = 10.0 * std::log10(num / den);
  // frames whose 512 samples lie inside clip 0's gap hold log10(0 + 1e-9) = -9 in every bin
  int silent = 0;
  for (long long t = 0; t < T; ++t) {
    const long long lo = t * hop - n_fft / 2 + (n_fft - win_length) / 2, hi = lo + win_length;
    if (lo >= gaps[0] && hi <= gaps[1]) silent += std::fabs(mag0[100 * T + t] + 9.0f) < 1e-6f;
  }
  std::printf("%s: %d clips x %lld samples -> [%d, %lld, %lld]; round-trip SNR %.1f dB; %d gap frames at -9\n", aip_version(), B, L, B, F,
              T, snr, silent);
  return (snr >= 100.0 && silent >= 10) ? 0 : 4;
}
