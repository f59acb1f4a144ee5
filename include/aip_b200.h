/* aip_b200.h -- C ABI of the B200-native spectrogram front-end / back-end.
 *
 * One shared object (libaip_b200.so, sm_100a SASS only) that replaces, for batches of clips that
 * are already resident in GPU memory, the numeric core of the reference's Python functions
 * (savage-hacker14/ml-audio-inpainting; file:line relative to the reference root):
 *
 *   aip_stft_fwd_f32        librosa.stft as called by utils.extract_spectrogram   utils.py:192-234
 *                           + the callers' epilogues: |S|, log10(|S|+eps)          models/CNNBLSTM/dataset.py:103-119
 *                             log1p(|S|**p), angle(S), frame masks                 models/GAN/dataset.py:121-152
 *                             spectrum-domain gap                                  models/model_eval.py:146-154
 *                           + time-domain gap zeroing before the transform         utils.py:141-142, :180-183, add_gaps.py:28-32
 *   aip_istft_f32           librosa.istft as called by utils.spectrogram_to_audio  utils.py:316-327
 *                           (complex input, or magnitude * exp(j*phase), dB / 10** / expm1 prologue)
 *   aip_istft_blend_f32     mask blend + 10** + phase reuse + istft               models/CNNBLSTM/model.py:108, models/model_eval.py:160-189
 *   aip_istft_handoff_f32   the same for either model family (+ peak normalisation) models/GAN/train.py:470-482, models/model_eval.py:118-140
 *   aip_griffinlim_f32      librosa.griffinlim (momentum 0.99)                     utils.py:328-332
 *   aip_griffinlim_c64_f32  librosa.griffinlim handed a COMPLEX "magnitude"        utils.py:328-332 via tests/utils_test.py:624-645
 *   aip_mel_project_f32     librosa.feature.melspectrogram's filter-bank contraction  utils.py:268-277
 *   aip_mel_inverse_f32     pinv(mel basis) @ mel (+ sqrt for power spectrograms)    utils.py:375-383
 *   aip_db_heuristic_f32    "max < 0 and mean < 0 => dB" test                      utils.py:313-314
 *   aip_gap_zero_f32        zero a sample range per clip                           utils.py:180-183, add_gaps.py:28-32, pre_process_dataset.py:38
 *   aip_gap_mask_f32        dense sample-domain 1/0 mask                           utils.py:141-142
 *   aip_frame_mask_f32      dense [F,T] frame mask                                 models/CNNBLSTM/dataset.py:115-118, models/GAN/dataset.py:150-152
 *   aip_peak_normalize_f32  librosa.util.normalize (norm=inf)                      utils.py:84
 *   aip_stft_gap_variants_f32  the gaps_per_audio loop of one dataset item        models/CNNBLSTM/dataset.py:93-111
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless stated otherwise;
 *   - the caller allocates every output and every workspace; nothing is allocated, freed or
 *     synchronised inside; work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - spectrogram layout is the reference's: [B, F = n_fft/2 + 1, T] with T contiguous
 *     (complex as interleaved float pairs); waveforms are [B, L] rows with an explicit pitch;
 *   - return value: 0 = ok, < 0 = argument / support error (AIP_ERR_*), > 0 = cudaError_t;
 *   - re-entrant and thread-safe.  The only mutable global state is the forward kernel's tile counter (dynamic tile
 *     schedule, 4 bytes of device memory): every (device, stream) pair owns one for good, it is zeroed on that stream right
 *     before each launch, and the zeroing + launch are enqueued under a lock -- launches on one stream are ordered by the
 *     stream, launches on different streams or devices never share a counter, whatever number of them is in flight
 *     (4096 distinct streams per device; one more returns cudaErrorLaunchOutOfResources);
 *   - AIP_* environment variables (experiment switches, csrc/aip_host.h) are read once, when the library is loaded;
 *   - there is NO CPU fallback: on a device that is not compute capability 10.0 (the library holds sm_100a code only)
 *     every entry point returns AIP_ERR_DEVICE.
 */
#ifndef AIP_B200_H_
#define AIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIP_OK 0
#define AIP_ERR_ARG (-1)          /* null pointer / negative size / inconsistent sizes */
#define AIP_ERR_UNSUPPORTED (-2)  /* parameter combination not implemented (e.g. n_fft not a power of two) */
#define AIP_ERR_DEVICE (-3)       /* current device is not sm_100 */
#define AIP_ERR_WORKSPACE (-4)    /* workspace too small */

/* forward magnitude epilogue */
#define AIP_MAG_NONE 0
#define AIP_MAG_ABS 1
#define AIP_MAG_LOG10_EPS 2       /* log10(|S| + eps) */
#define AIP_MAG_LOG1P_POW 3       /* log1p(|S| ** power) */
#define AIP_MAG_POW 4             /* |S| ** power */

/* inverse magnitude prologue */
#define AIP_DOM_LINEAR 0
#define AIP_DOM_POW10 1           /* 10 ** x */
#define AIP_DOM_DB 2              /* 10 ** (x / 20) */
#define AIP_DOM_EXPM1 3           /* expm1(x) */

typedef struct aip_stft_desc {
  int32_t n_fft;        /* power of two, 32 .. 4096; 512 runs the register-FFT kernels */
  int32_t hop;          /* > 0 */
  int32_t center;       /* 0: no padding; 1: zero-pad n_fft/2 on both sides (librosa >= 0.10 pad_mode="constant");
                           2: the same amount mirrored about the first / last sample (pad_mode="reflect", librosa < 0.10's
                           default; forward only -- the inverse trims n_fft/2 either way) */
  int32_t win_length;   /* 0 or n_fft: every tap may be non-zero.  Otherwise the caller GUARANTEES that `window` is a
                           win_length window centre-padded with zeros (librosa util.pad_center, what librosa.stft /
                           istft build from win_length): the kernels then skip the zero taps (config.py: 384 in 512) */
  const float* window;  /* device [n_fft]: the centre-padded fft window (librosa.filters.get_window + pad_center) */
} aip_stft_desc;

/* Number of frames librosa.stft yields for L samples; < 0 on error. */
int64_t aip_num_frames(int64_t L, int32_t n_fft, int32_t hop, int32_t center);
/* Output length of librosa.istft for T frames (length == 0: natural length). */
int64_t aip_istft_length(int64_t T, int32_t n_fft, int32_t hop, int32_t center, int64_t length);

/* Forward transform + fused epilogues.  Any of the four outputs may be null (mag_out must be
 * non-null iff mag_kind != AIP_MAG_NONE).  T_out <= aip_num_frames(...) frames are written.
 *   gap_samples  [B,2] int32 or null: samples [g0,g1) of clip b are treated as zero;
 *   zero_frames  [B,2] int32 or null: the spectrum of frames [f0,f1) is zeroed before the epilogue;
 *   mask_frames  [B,2] int32 or null: frame range for mask_out; mask_in_gap_is_one selects the
 *                CNNBLSTM (1 inside, 0 outside) or GAN (0 inside, 1 outside) convention.          */
int aip_stft_fwd_f32(const aip_stft_desc* desc,
                     const float* wave, int64_t B, int64_t L, int64_t wave_pitch,
                     const int32_t* gap_samples, const int32_t* zero_frames,
                     const int32_t* mask_frames, int32_t mask_in_gap_is_one,
                     int32_t mag_kind, float eps, float power, int64_t T_out,
                     float* spec_out /* [B,F,T_out,2] */, float* mag_out, float* phase_out, float* mask_out,
                     void* stream);

/* Inverse transform.  Input: spec (complex) or mag (+ optional phase).  `T` is the frame count
 * (row length) of the inputs; `length` = 0 for librosa's natural length hop*(T-1) (center) else the
 * requested length.  inv_wss: device [out_len] from aip_inv_window_sumsquare_f32.
 * db_flags: [B] int32 or null; non-zero entries force the dB prologue for that clip.
 * workspace: only needed off the fused n_fft = 512 path (aip_istft_workspace_bytes).              */
int aip_istft_f32(const aip_stft_desc* desc,
                  const float* spec, const float* mag, const float* phase, int32_t mag_domain,
                  const int32_t* db_flags,
                  int64_t B, int64_t T, int64_t length, const float* inv_wss,
                  float* wave_out, int64_t out_pitch,
                  void* workspace, size_t workspace_bytes, void* stream);
/* The model hand-off of the CNN-BLSTM path fused into the inverse prologue (reference models/CNNBLSTM/model.py:108,
 * models/model_eval.py:160-163, :179-189):  m = model_out * blend_mask + blend_in * (1 - blend_mask);
 * magnitude = 10 ** m (AIP_DOM_POW10) or 10 ** (m / 20) (AIP_DOM_DB);  X = magnitude * exp(j * phase);  istft(X).
 * All four inputs are [B,F,T] float32.                                                                        */
int aip_istft_blend_f32(const aip_stft_desc* desc, const float* model_out, const float* blend_in,
                        const float* blend_mask, const float* phase, int32_t mag_domain,
                        int64_t B, int64_t T, int64_t length, const float* inv_wss,
                        float* wave_out, int64_t out_pitch,
                        void* workspace, size_t workspace_bytes, void* stream);
/* The model hand-off in general form: magnitudes blended with a mask in either family's convention, any magnitude domain,
 * phase reuse, istft and -- when `peaks` is non-null -- the peak normalisation of save_audio (utils.py:84) on top.
 *   mask_keeps_input = 0   m = model_out * mask + blend_in * (1 - mask)      mask is 1 INSIDE the gap
 *                          (StackedBLSTMCNN.reconstruct_spectrogram, models/CNNBLSTM/model.py:108)
 *   mask_keeps_input = 1   m = model_out * (1 - mask) + blend_in * mask      mask is 1 OUTSIDE the gap
 *                          (combined_log_mag, models/GAN/train.py:473)
 *   mag_domain: AIP_DOM_LINEAR (GAN/train.py:476 hands the log1p-domain blend to spectrogram_to_audio as it is),
 *               AIP_DOM_POW10 (models/model_eval.py:163), AIP_DOM_DB, AIP_DOM_EXPM1 (undoing the GAN's log1p).
 *   peaks: device [B] float or null; pcm_out: device [B, pcm_pitch] int16 or null -- see aip_istft_normalized_f32
 *   (peaks null + pcm_out: the un-normalised waveform is quantised as it is).                                           */
int aip_istft_handoff_f32(const aip_stft_desc* desc, const float* model_out, const float* blend_in,
                          const float* blend_mask, int32_t mask_keeps_input, const float* phase, int32_t mag_domain,
                          int64_t B, int64_t T, int64_t length, const float* inv_wss,
                          float* wave_out, int64_t out_pitch, float* peaks, int16_t* pcm_out, int64_t pcm_pitch,
                          void* workspace, size_t workspace_bytes, void* stream);
/* 0 when the (n_fft, hop, center) combination runs the fused n_fft = 512 kernel. */
size_t aip_istft_workspace_bytes(const aip_stft_desc* desc, int64_t B, int64_t T);

/* 1 / window_sumsquare (where > FLT_MIN, else 1), accumulated in float32 frame by frame exactly as
 * librosa.filters.window_sumsquare does; written to inv_wss[out_len].                              */
int aip_inv_window_sumsquare_f32(const aip_stft_desc* desc, int64_t T, int64_t length,
                                 float* inv_wss, int64_t out_len, void* stream);

/* Griffin-Lim: n_iter x (istft, stft, phase update) and a final istft.
 *   mag     [B,F,T] linear magnitudes (8-byte aligned);
 *   angles  [B,F,T,2] in: initial unit phasors (caller-drawn); used as the iteration state (16-byte aligned);
 *   tprev   [B,F,T,2] scratch (16-byte aligned);  wave_out [B, out_len] doubles as the iteration buffer;
 *   workspace: aip_istft_workspace_bytes(...) rounded up to 16, optionally followed by B*F*T*8 more bytes -- with that
 *   extra array the rebuilt spectra of consecutive iterations ping-pong and no `tprev = rebuilt` copy is made.       */
/* librosa.griffinlim(init="random") (utils.py:330-332 leaves librosa's default): n unit phasors exp(2 pi j u), u uniform in
 * [0, 1), from a Philox4x32-10 counter stream keyed by `seed` -- the `angles` array aip_griffinlim_f32 starts from.       */
int aip_random_phasors_f32(float* angles /* [n, 2] */, int64_t n, uint64_t seed, void* stream);
int aip_griffinlim_f32(const aip_stft_desc* desc, const float* mag, float* angles, float* tprev,
                       int64_t B, int64_t T, int32_t n_iter, float momentum, const float* inv_wss,
                       float* wave_out, int64_t out_pitch,
                       void* workspace, size_t workspace_bytes, void* stream);

/* librosa.griffinlim when the caller hands it a COMPLEX matrix as "magnitude" (tests/utils_test.py:624-645 pass
 * extract_spectrogram's complex output straight in): librosa multiplies the unit phasors by S as it is -- a complex product.
 *   spec    [B,F,T,2] the complex "magnitude" S;  angles / tprev / workspace as for aip_griffinlim_f32, and the workspace MUST
 *   hold the extra B*F*T*8 bytes (the rebuilt spectra ping-pong; the update runs as its own kernel).                     */
int aip_griffinlim_c64_f32(const aip_stft_desc* desc, const float* spec, float* angles, float* tprev,
                           int64_t B, int64_t T, int32_t n_iter, float momentum, const float* inv_wss,
                           float* wave_out, int64_t out_pitch,
                           void* workspace, size_t workspace_bytes, void* stream);

/* mel_out[b, m, t] = sum_f basis[m, f] * spec_pow[b, f, t]: the contraction librosa.feature.melspectrogram applies to
 * |stft| ** power (utils.py:268-277; spec_pow is aip_stft_fwd_f32's AIP_MAG_POW / AIP_MAG_ABS output).
 *   basis  [n_mels, F] row major (librosa.filters.mel: triangular, so each row is non-zero on one short bin range);
 *   bands  [n_mels, 2] int32: per row a range [f0, f1) that contains every non-zero entry of that row (host-computed).    */
int aip_mel_project_f32(const float* basis, const int32_t* bands, const float* spec_pow, int64_t B, int64_t F, int64_t T,
                        int64_t n_mels, float* mel_out, void* stream);
/* out[b, f, t] = sum_m inv_basis[f, m] * mel[b, m, t], then sqrt() when take_sqrt != 0 (utils.py:375-383: pinv of the mel
 * basis applied to a mel spectrogram).  Negative projections become NaN under the square root exactly as np.sqrt makes
 * them, EXCEPT those below the noise floor of an fp32 power spectrogram (|v| <= 1e-9 x the clip's largest projection), which
 * are rounding residue of a non-negative exact value and are taken as 0.  peaks: [B] float scratch (needed for take_sqrt). */
int aip_mel_inverse_f32(const float* inv_basis, const float* mel, int64_t B, int64_t F, int64_t T, int64_t n_mels,
                        int32_t take_sqrt, float* out, float* peaks, void* stream);

/* flags[b] = (max(x_b) < 0 && mean(x_b) < 0), x_b = x[b*n .. (b+1)*n). */
int aip_db_heuristic_f32(const float* x, int64_t B, int64_t n, int32_t* flags, void* stream);

/* out[b, s] = (g0_b <= s < g1_b) ? 0 : in[b, s]   (in == out allowed). */
int aip_gap_zero_f32(const float* in, int64_t in_pitch, float* out, int64_t out_pitch,
                     int64_t B, int64_t L, const int32_t* gap_samples, void* stream);
/* mask[b, s] = (g0_b <= s < g1_b) ? 0 : 1 */
int aip_gap_mask_f32(float* mask, int64_t pitch, int64_t B, int64_t L, const int32_t* gap_samples, void* stream);
/* mask[b, f, t] = 1/0 by frame range (see aip_stft_fwd_f32). */
int aip_frame_mask_f32(float* mask, int64_t B, int64_t F, int64_t T, const int32_t* mask_frames,
                       int32_t mask_in_gap_is_one, void* stream);
/* aip_istft_f32 followed by librosa.util.normalize (norm = inf) of every clip -- what the reference's callers do with
 * spectrogram_to_audio's result before writing it (utils.save_audio, utils.py:84; models/CNNBLSTM/train.py:184-186,
 * models/model_eval.py:130,179).  The per-clip peak max|y| is taken inside the overlap-add of the inverse kernel (one
 * atomic per warp and tile), so the waveform is read back only once, by the in-place scaling pass.
 * peaks: device [B] float, out (the peak of each clip BEFORE scaling); clips whose peak is < FLT_MIN stay unscaled.
 * pcm_out: null, or device [B, pcm_pitch] int16 -- then that one pass writes what save_audio puts into its 16-bit FLAC
 * (utils.py:87, soundfile's float -> PCM_16: x * 32768, round half to even, clip to [-32768, 32767]; SURVEY 8f rank 1
 * "optional int16 quantisation for the encoder") and wave_out keeps the UN-normalised waveform: the float scaling pass
 * and half of the device -> host bytes are saved.                                                                    */
int aip_istft_normalized_f32(const aip_stft_desc* desc,
                             const float* spec, const float* mag, const float* phase, int32_t mag_domain,
                             const int32_t* db_flags,
                             int64_t B, int64_t T, int64_t length, const float* inv_wss,
                             float* wave_out, int64_t out_pitch, float* peaks, int16_t* pcm_out, int64_t pcm_pitch,
                             void* workspace, size_t workspace_bytes, void* stream);
/* Gap variants of one file: models/CNNBLSTM/dataset.py:93-111 loads a file gaps_per_audio times, zeroes a
 * different random range each time (utils.add_random_gap, utils.py:179-183) and runs a full STFT per gap; a gap only changes
 * the frames whose n_fft-sample span meets it.  Variant v = i * G + j is gap j of wave row i:
 *   mag_out[v] := clean_mag[i] everywhere, then the <= ceil((gap_len_max + n_fft) / hop) + 1 frames around gap_samples[v]
 *   are re-transformed from wave row i with samples [gap_samples[v][0], gap_samples[v][1]) zeroed (same kernel, same
 *   arithmetic as aip_stft_fwd_f32 with gap_samples: the result is bit-identical to G full transforms).
 * clean_mag [N, F, T_out] is the caller's aip_stft_fwd_f32(..., mag_kind, eps, T_out) of the un-gapped rows.
 * gap_len_max: the caller's bound on gap_samples[v][1] - gap_samples[v][0] (a longer gap is NOT detected).
 * mag_kind: AIP_MAG_ABS | AIP_MAG_LOG10_EPS | AIP_MAG_LOG1P_POW (power 1); n_fft 512 only (else AIP_ERR_UNSUPPORTED).
 * workspace: 16-byte aligned device scratch of aip_stft_gap_variants_workspace_bytes(N, G) bytes (per-variant tile metadata). */
size_t aip_stft_gap_variants_workspace_bytes(int64_t N, int64_t G);
int aip_stft_gap_variants_f32(const aip_stft_desc* desc, const float* wave, int64_t N, int64_t L, int64_t wave_pitch,
                              int64_t G, const int32_t* gap_samples /* [N*G, 2] */, int32_t gap_len_max,
                              int32_t mag_kind, float eps, int64_t T_out, const float* clean_mag /* [N, F, T_out] */,
                              float* mag_out /* [N*G, F, T_out] */, void* workspace, size_t workspace_bytes, void* stream);

/* y_b / max|y_b| unless max|y_b| < FLT_MIN; peaks: [B] float scratch/out (the per-clip max|y|). */
/* save_audio's tail (utils.py:83-87) for a waveform that is already on the device: librosa.util.normalize + soundfile's
 * float -> PCM_16 conversion (see aip_istft_normalized_f32), one read of `in`, int16 out.
 *   AIP_PCM_RAW          quantise as is (save_audio(normalize=False)); peaks unused
 *   AIP_PCM_NORMALIZE    take each clip's peak here (one more read of `in`), write it to peaks[B], divide, quantise
 *   AIP_PCM_PEAKS_GIVEN  peaks[B] already holds max|in| per clip                                                        */
#define AIP_PCM_RAW 0
#define AIP_PCM_NORMALIZE 1
#define AIP_PCM_PEAKS_GIVEN 2
int aip_wave_to_pcm16_f32(const float* in, int64_t in_pitch, int16_t* pcm, int64_t pcm_pitch, int64_t B, int64_t L,
                          int32_t peaks_mode, float* peaks, void* stream);
int aip_peak_normalize_f32(const float* in, int64_t in_pitch, float* out, int64_t out_pitch,
                           int64_t B, int64_t L, float* peaks, void* stream);

/* Re-read the AIP_* experiment switches from the environment (they are otherwise read once, at load).  For A/B timing
 * scripts and the tests of alternative kernels; not thread-safe against concurrent launches.                         */
void aip_debug_reload_env(void);
const char* aip_status_string(int status);
/* "aip_b200 <version> sm_100a" */
const char* aip_version(void);
/* 1 when the current device can run the kernels (compute capability 10.0), else 0. */
int aip_device_supported(void);

#ifdef __cplusplus
}
#endif
#endif /* AIP_B200_H_ */
