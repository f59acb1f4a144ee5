/* aip_codec.h -- host-side FLAC codec of the drop-in's audio file I/O (plain C, no CUDA; lib/libaip_codec.so).
 *
 * Replaces, for the containers the reference's data uses (LibriSpeech: 16 kHz / mono / 16-bit FLAC), what the reference reaches
 * through third-party libraries:
 *   decode   utils.load_audio -> librosa.load -> soundfile / libsndfile                       /root/reference/utils.py:36
 *   encode   utils.save_audio -> soundfile.write(..., format='flac') (PCM_16)                 /root/reference/utils.py:87
 *            add_gaps.insert_gap -> sf.write                                                  /root/reference/add_gaps.py:36
 * The arithmetic around it (int -> float scaling, padding / truncation, peak normalisation, float -> PCM_16) stays where it was:
 * ml_audio_inpainting_b200/audio_io.py and the device kernels of aip_b200.h.
 *
 * Conventions: caller-owned buffers, no allocation visible to the caller, no global state besides two CRC tables,
 * re-entrant (decode different files from different threads), return value < 0 = AIP_CODEC_ERR_*.
 */
#ifndef AIP_CODEC_H_
#define AIP_CODEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIP_CODEC_ERR_ARG (-1)
#define AIP_CODEC_ERR_FORMAT (-2)
#define AIP_CODEC_ERR_TRUNCATED (-3)
#define AIP_CODEC_ERR_SYNC (-4)
#define AIP_CODEC_ERR_CAPACITY (-5)

typedef struct aip_flac_info {
  int32_t sample_rate;
  int32_t channels;
  int32_t bits_per_sample;
  int32_t min_blocksize;
  int32_t max_blocksize;
  int64_t total_samples;        /* per channel; 0 = unknown */
  uint8_t md5[16];              /* MD5 of the little-endian interleaved PCM; all zero = not recorded */
} aip_flac_info;

/* STREAMINFO of a FLAC byte string. */
int aip_flac_info_read(const uint8_t* data, size_t n, aip_flac_info* info);

/* Decode into out[sample * channels + channel] (int32, capacity cap_samples per channel).  max_samples > 0 stops after the frame
 * that reaches that many samples per channel (so up to max_blocksize - 1 more may be returned); 0 decodes everything.
 * Returns the samples per channel written, or an error.  info may be null. */
int64_t aip_flac_decode(const uint8_t* data, size_t n, int64_t max_samples, int32_t* out, int64_t cap_samples,
                        aip_flac_info* info);

/* Encode 16-bit PCM pcm[sample * channels + channel] as a FLAC stream with frames of `blocksize` samples.  md5: the caller's
 * MD5 of the little-endian PCM bytes (STREAMINFO signature).  Returns the bytes written to out (capacity cap; 42 + 2.2 bytes per
 * sample and channel + 32 per frame always suffices), or an error. */
int64_t aip_flac_encode16(const int16_t* pcm, int64_t n, int32_t channels, int32_t sample_rate, int32_t blocksize,
                          const uint8_t md5[16], uint8_t* out, size_t cap);

const char* aip_codec_status_string(int status);

#ifdef __cplusplus
}
#endif
#endif  /* AIP_CODEC_H_ */
