/* Host-side FLAC codec behind utils.load_audio / utils.save_audio (reference utils.py:36 librosa.load -> libsndfile,
 * utils.py:87 soundfile.write; SURVEY 8f rank 4 "host-side I/O").  Not part of the GPU path and no CUDA in here: plain C,
 * built with gcc into lib/libaip_codec.so, called through ctypes (which releases the GIL: files decode in parallel from a
 * Python thread pool).  Declarations and the entry points' contracts: include/aip_codec.h.
 *
 *   decode  every subframe type (CONSTANT / VERBATIM / FIXED 0..4 / LPC 1..32), Rice partitions with 4- or 5-bit parameters and
 *           escape codes, wasted bits, independent / left-side / side-right / mid-side channels, 4..32 bits per sample,
 *           fixed or variable block size; stops after the frame that reaches `max_samples` (load_audio needs the first
 *           sample_rate * max_len samples only).  Frame CRCs are not checked (the STREAMINFO MD5 is, by the caller, on request).
 *   encode  16-bit, independent channels, FIXED predictor of the order with the smallest sum |residual|, one Rice partition --
 *           bit for bit the stream the pure-Python checker codec writes (tests/test_codec.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aip_codec.h"

/* ------------------------------------------------------------------------------------------------ bit reader */
typedef struct {
  const uint8_t* p;
  size_t n;        /* bytes */
  size_t pos;      /* bit position */
  int err;
} bitr;

static inline uint64_t be64(const uint8_t* p) {
  uint64_t w;
  memcpy(&w, p, 8);
  return __builtin_bswap64(w);
}

/* n <= 32 bits, big endian */
static inline uint32_t br_read(bitr* b, int n) {
  if (n == 0) return 0;
  if (b->pos + (size_t)n > b->n * 8) { b->err = 1; b->pos = b->n * 8; return 0; }
  const size_t byte = b->pos >> 3;
  const int off = (int)(b->pos & 7);
  b->pos += (size_t)n;
  if (byte + 8 <= b->n) return (uint32_t)((be64(b->p + byte) << off) >> (64 - n));      /* off + n <= 39 */
  uint64_t acc = 0;
  const int need = (off + n + 7) >> 3;          /* <= 5 bytes, all inside the buffer */
  for (int i = 0; i < need; ++i) acc = (acc << 8) | b->p[byte + i];
  acc >>= (need * 8 - off - n);
  return (uint32_t)(acc & (n == 32 ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

static inline int32_t br_signed(bitr* b, int n) {
  if (n == 0) return 0;
  const uint32_t v = br_read(b, n);
  if (n == 32) return (int32_t)v;
  return (v >> (n - 1)) ? (int32_t)v - (int32_t)(1u << n) : (int32_t)v;
}

/* number of zero bits before the next one bit (which is consumed) */
static inline uint32_t br_unary(bitr* b) {
  uint32_t z = 0;
  const size_t end = b->n * 8;
  while (b->pos < end) {
    const size_t byte = b->pos >> 3;
    const int off = (int)(b->pos & 7);
    if (byte + 8 <= b->n) {
      const uint64_t w = be64(b->p + byte) << off;                 /* 64 - off valid bits, left aligned */
      if (w) {
        const int lead = __builtin_clzll(w);
        b->pos += (size_t)lead + 1;
        return z + (uint32_t)lead;
      }
      z += (uint32_t)(64 - off);
      b->pos += (size_t)(64 - off);
      continue;
    }
    const uint8_t rest = (uint8_t)(b->p[byte] << off);             /* the last few bytes: one at a time */
    if (rest) {
      const int lead = __builtin_clz((uint32_t)rest) - 24;
      b->pos += (size_t)lead + 1;
      return z + (uint32_t)lead;
    }
    z += (uint32_t)(8 - off);
    b->pos += (size_t)(8 - off);
  }
  b->err = 1;
  return z;
}

static uint64_t br_utf8(bitr* b) {
  uint32_t x = br_read(b, 8);
  if (x < 0x80) return x;
  int n = 0;
  while (x & (0x80u >> n)) ++n;
  if (n < 2 || n > 7) { b->err = 1; return 0; }
  uint64_t v = x & ((1u << (7 - n)) - 1);
  for (int i = 0; i < n - 1; ++i) v = (v << 6) | (br_read(b, 8) & 0x3F);
  return v;
}

/* ------------------------------------------------------------------------------------------------ decode */
static const int kBlockSizes[16] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
static const int kSampleSizes[8] = {0, 8, 12, 0, 16, 20, 24, 32};

static int read_residual(bitr* b, int blocksize, int order, int64_t* out) {
  const uint32_t method = br_read(b, 2);
  if (method > 1) return AIP_CODEC_ERR_FORMAT;
  const int pbits = method == 0 ? 4 : 5;
  const uint32_t escape = (1u << pbits) - 1;
  const int porder = (int)br_read(b, 4);
  const int nparts = 1 << porder;
  if ((blocksize >> porder) << porder != blocksize && porder > 0) return AIP_CODEC_ERR_FORMAT;
  int w = 0;
  for (int part = 0; part < nparts; ++part) {
    int n = (blocksize >> porder) - (part == 0 ? order : 0);
    if (n < 0 || w + n > blocksize - order) return AIP_CODEC_ERR_FORMAT;
    const uint32_t k = br_read(b, pbits);
    if (k == escape) {
      const int raw = (int)br_read(b, 5);
      for (int i = 0; i < n; ++i) out[w + i] = raw ? br_signed(b, raw) : 0;
    } else {
      /* one 8-byte window per residual: unary run + stop bit + k low bits fit its >= 57 valid bits almost always; the rare
         long run (or the last 8 bytes of the stream) takes the general readers */
      size_t pos = b->pos;
      const size_t safe_bits = b->n >= 8 ? (b->n - 8) * 8 : 0;
      for (int i = 0; i < n; ++i) {
        uint64_t u;
        if (pos <= safe_bits) {
          const uint64_t win = be64(b->p + (pos >> 3)) << (pos & 7);
          const int lead = win ? __builtin_clzll(win) : 64;
          if (lead + 1 + (int)k <= 57) {
            u = (uint64_t)lead;
            if (k) u = (u << k) | ((win << (lead + 1)) >> (64 - k));
            pos += (size_t)lead + 1 + k;
            out[w + i] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
            continue;
          }
        }
        b->pos = pos;
        u = br_unary(b);
        if (k) u = (u << k) | br_read(b, (int)k);
        pos = b->pos;
        out[w + i] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
      }
      b->pos = pos;
    }
    if (b->err) return AIP_CODEC_ERR_TRUNCATED;
    w += n;
  }
  return w == blocksize - order ? 0 : AIP_CODEC_ERR_FORMAT;
}

/* one subframe -> s[0..blocksize) (int64 because side channels of 32-bit streams need 33 bits) */
static int read_subframe(bitr* b, int blocksize, int bps, int64_t* s, int64_t* res) {
  static const int fixed[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
  if (br_read(b, 1)) return AIP_CODEC_ERR_FORMAT;
  const int typ = (int)br_read(b, 6);
  int wasted = 0;
  if (br_read(b, 1)) {
    wasted = (int)br_unary(b) + 1;
    bps -= wasted;
  }
  if (bps < 1 || bps > 33) return AIP_CODEC_ERR_FORMAT;
  int rc = 0;
  if (typ == 0) {
    const int64_t v = bps <= 32 ? br_signed(b, bps) : 0;
    for (int i = 0; i < blocksize; ++i) s[i] = v;
  } else if (typ == 1) {
    for (int i = 0; i < blocksize; ++i) {
      if (bps <= 32) s[i] = br_signed(b, bps);
      else { const int64_t hi = br_signed(b, 1); s[i] = hi * (1LL << 32) + br_read(b, 32); }
    }
  } else if ((typ >= 8 && typ <= 12) || typ >= 32) {
    const int order = typ >= 32 ? (typ & 31) + 1 : typ - 8;
    if (order > blocksize) return AIP_CODEC_ERR_FORMAT;
    for (int i = 0; i < order; ++i) {
      if (bps <= 32) s[i] = br_signed(b, bps);
      else { const int64_t hi = br_signed(b, 1); s[i] = hi * (1LL << 32) + br_read(b, 32); }
    }
    int64_t c[32];
    int shift = 0;
    if (typ >= 32) {
      const int precision = (int)br_read(b, 4) + 1;
      if (precision == 16) return AIP_CODEC_ERR_FORMAT;
      shift = br_signed(b, 5);
      if (shift < 0) return AIP_CODEC_ERR_FORMAT;
      for (int i = 0; i < order; ++i) c[i] = br_signed(b, precision);
    } else {
      for (int i = 0; i < order; ++i) c[i] = fixed[order][i];
    }
    rc = read_residual(b, blocksize, order, res);
    if (rc) return rc;
    if (order == 0) {
      memcpy(s, res, (size_t)blocksize * sizeof(int64_t));
    } else {
      /* s[n] = res[n - order] + (sum_i c[i] s[n - 1 - i] >> shift), arithmetic shift = floor as the format specifies.  The
         orders encoders actually emit (libFLAC -5 .. -8: up to 12) get a fully unrolled inner loop: with the trip count a
         run-time value the restore took most of the decode time */
#define AIP_LPC_CASE(ORD)                                                        \
  case ORD:                                                                      \
    for (int n = ORD; n < blocksize; ++n) {                                      \
      int64_t acc = 0;                                                           \
      /* the term that depends on the sample just restored goes last: the others are off the critical path */ \
      _Pragma("GCC unroll 16") for (int i = ORD - 1; i >= 0; --i) acc += c[i] * s[n - 1 - i]; \
      s[n] = res[n - ORD] + (acc >> shift);                                      \
    }                                                                            \
    break;
      switch (order) {
        AIP_LPC_CASE(1) AIP_LPC_CASE(2) AIP_LPC_CASE(3) AIP_LPC_CASE(4) AIP_LPC_CASE(5) AIP_LPC_CASE(6) AIP_LPC_CASE(7)
        AIP_LPC_CASE(8) AIP_LPC_CASE(9) AIP_LPC_CASE(10) AIP_LPC_CASE(11) AIP_LPC_CASE(12)
        default:
          for (int n = order; n < blocksize; ++n) {
            int64_t acc = 0;
            for (int i = 0; i < order; ++i) acc += c[i] * s[n - 1 - i];
            s[n] = res[n - order] + (acc >> shift);
          }
      }
#undef AIP_LPC_CASE
    }
  } else {
    return AIP_CODEC_ERR_FORMAT;
  }
  if (b->err) return AIP_CODEC_ERR_TRUNCATED;
  if (wasted)
    for (int i = 0; i < blocksize; ++i) s[i] = s[i] * (1LL << wasted);
  return 0;
}

static int parse_streaminfo(const uint8_t* data, size_t n, aip_flac_info* info, size_t* audio_start) {
  if (n < 8 || memcmp(data, "fLaC", 4) != 0) return AIP_CODEC_ERR_FORMAT;
  size_t pos = 4;
  int have = 0;
  for (;;) {
    if (pos + 4 > n) return AIP_CODEC_ERR_TRUNCATED;
    const int last = data[pos] >> 7, btype = data[pos] & 0x7F;
    const size_t len = ((size_t)data[pos + 1] << 16) | ((size_t)data[pos + 2] << 8) | data[pos + 3];
    const uint8_t* body = data + pos + 4;
    if (pos + 4 + len > n) return AIP_CODEC_ERR_TRUNCATED;
    if (btype == 0) {
      if (len < 34) return AIP_CODEC_ERR_FORMAT;
      info->min_blocksize = (body[0] << 8) | body[1];
      info->max_blocksize = (body[2] << 8) | body[3];
      uint64_t x = 0;
      for (int i = 10; i < 18; ++i) x = (x << 8) | body[i];
      info->sample_rate = (int32_t)(x >> 44);
      info->channels = (int32_t)((x >> 41) & 7) + 1;
      info->bits_per_sample = (int32_t)((x >> 36) & 31) + 1;
      info->total_samples = (int64_t)(x & ((1ull << 36) - 1));
      memcpy(info->md5, body + 18, 16);
      have = 1;
    }
    pos += 4 + len;
    if (last) break;
  }
  if (!have) return AIP_CODEC_ERR_FORMAT;
  *audio_start = pos;
  return 0;
}

int aip_flac_info_read(const uint8_t* data, size_t n, aip_flac_info* info) {
  size_t start = 0;
  if (!data || !info) return AIP_CODEC_ERR_ARG;
  return parse_streaminfo(data, n, info, &start);
}

int64_t aip_flac_decode(const uint8_t* data, size_t n, int64_t max_samples, int32_t* out, int64_t cap_samples,
                        aip_flac_info* info_out) {
  aip_flac_info info;
  size_t start = 0;
  if (!data || !out || cap_samples < 0) return AIP_CODEC_ERR_ARG;
  int rc = parse_streaminfo(data, n, &info, &start);
  if (rc) return rc;
  if (info_out) *info_out = info;
  const int nch = info.channels;
  bitr b = {data, n, start * 8, 0};
  int64_t got = 0;
  const int64_t want = info.total_samples ? info.total_samples : -1;
  int64_t* buf = NULL;
  int cap_block = 0;
  while (b.pos + 16 <= n * 8 && (want < 0 || got < want)) {
    if (max_samples > 0 && got >= max_samples) break;
    if (br_read(&b, 14) != 0x3FFE) { rc = AIP_CODEC_ERR_SYNC; break; }
    br_read(&b, 2);                                     /* reserved, blocking strategy (the number is parsed either way) */
    const int bs_code = (int)br_read(&b, 4), sr_code = (int)br_read(&b, 4), ch_code = (int)br_read(&b, 4);
    const int ss_code = (int)br_read(&b, 3);
    br_read(&b, 1);
    br_utf8(&b);
    int blocksize;
    if (bs_code == 6) blocksize = (int)br_read(&b, 8) + 1;
    else if (bs_code == 7) blocksize = (int)br_read(&b, 16) + 1;
    else blocksize = kBlockSizes[bs_code];
    if (blocksize <= 0) { rc = AIP_CODEC_ERR_FORMAT; break; }
    if (sr_code == 12) br_read(&b, 8);
    else if (sr_code == 13 || sr_code == 14) br_read(&b, 16);
    br_read(&b, 8);                                     /* CRC-8 */
    int bps = ss_code ? kSampleSizes[ss_code] : info.bits_per_sample;
    if (bps == 0) bps = info.bits_per_sample;
    if (b.err) { rc = AIP_CODEC_ERR_TRUNCATED; break; }
    if (blocksize > cap_block) {
      free(buf);
      buf = (int64_t*)malloc((size_t)blocksize * 3 * sizeof(int64_t));
      if (!buf) { rc = AIP_CODEC_ERR_ARG; break; }
      cap_block = blocksize;
    }
    int64_t *c0 = buf, *c1 = buf + cap_block, *res = buf + 2 * cap_block;
    int64_t keep = blocksize;
    if (want >= 0 && got + keep > want) keep = want - got;
    if (got + keep > cap_samples) { rc = AIP_CODEC_ERR_CAPACITY; break; }
    if (ch_code < 8) {
      if (ch_code + 1 != nch) { rc = AIP_CODEC_ERR_FORMAT; break; }
      for (int c = 0; c < nch && !rc; ++c) {
        rc = read_subframe(&b, blocksize, bps, c0, res);
        if (!rc)
          for (int64_t i = 0; i < keep; ++i) out[(got + i) * nch + c] = (int32_t)c0[i];
      }
      if (rc) break;
    } else if (ch_code <= 10) {
      if (nch != 2) { rc = AIP_CODEC_ERR_FORMAT; break; }
      const int b0 = ch_code == 9 ? bps + 1 : bps, b1 = ch_code == 9 ? bps : bps + 1;
      rc = read_subframe(&b, blocksize, b0, c0, res);
      if (!rc) rc = read_subframe(&b, blocksize, b1, c1, res);
      if (rc) break;
      for (int64_t i = 0; i < keep; ++i) {
        int64_t l, r;
        if (ch_code == 8) { l = c0[i]; r = c0[i] - c1[i]; }                       /* left / side */
        else if (ch_code == 9) { l = c1[i] + c0[i]; r = c1[i]; }                  /* side / right */
        else { const int64_t mid = (c0[i] * 2) | (c1[i] & 1); l = (mid + c1[i]) >> 1; r = (mid - c1[i]) >> 1; }
        out[(got + i) * 2] = (int32_t)l;
        out[(got + i) * 2 + 1] = (int32_t)r;
      }
    } else {
      rc = AIP_CODEC_ERR_FORMAT;
      break;
    }
    b.pos = (b.pos + 7) & ~(size_t)7;
    br_read(&b, 16);                                    /* CRC-16 */
    if (b.err) { rc = AIP_CODEC_ERR_TRUNCATED; break; }
    got += keep;
  }
  free(buf);
  return rc ? rc : got;
}

/* ------------------------------------------------------------------------------------------------ encode */
typedef struct {
  uint8_t* p;
  size_t cap, pos;   /* bytes written */
  uint64_t acc;      /* pending bits, right aligned */
  int nacc;          /* < 8 between calls */
  int err;
} bitw;

static inline void bw_put(bitw* w, uint64_t v, int n) {          /* n <= 32 */
  if (n == 0) return;
  w->acc = (w->acc << n) | (v & (n == 32 ? 0xFFFFFFFFull : ((1ull << n) - 1)));
  w->nacc += n;
  while (w->nacc >= 8) {
    if (w->pos >= w->cap) { w->err = 1; w->nacc = 0; return; }
    w->p[w->pos++] = (uint8_t)(w->acc >> (w->nacc - 8));
    w->nacc -= 8;
  }
}
static inline void bw_zeros_then_one(bitw* w, uint64_t zeros) {
  while (zeros >= 32) { bw_put(w, 0, 32); zeros -= 32; }
  bw_put(w, 1, (int)zeros + 1);
}
static inline void bw_flush(bitw* w) {                            /* zero-pad to a byte boundary */
  if (w->nacc) bw_put(w, 0, 8 - w->nacc);
}

static uint8_t crc8_tab[256];
static uint16_t crc16_tab[256];
/* filled when the library is loaded: no lazily-initialised state, the entry points are re-entrant */
__attribute__((constructor)) static void crc_init(void) {
  for (int i = 0; i < 256; ++i) {
    uint8_t c = (uint8_t)i;
    for (int k = 0; k < 8; ++k) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
    crc8_tab[i] = c;
    uint16_t d = (uint16_t)(i << 8);
    for (int k = 0; k < 8; ++k) d = (d & 0x8000) ? (uint16_t)((d << 1) ^ 0x8005) : (uint16_t)(d << 1);
    crc16_tab[i] = d;
  }
}

static void encode_subframe(bitw* w, const int16_t* pcm, int n, int stride, int64_t* x, int64_t* r) {
  for (int i = 0; i < n; ++i) x[i] = pcm[(size_t)i * stride];
  int constant = n > 0;
  for (int i = 1; i < n && constant; ++i) constant = x[i] == x[0];
  if (constant) {                                                 /* CONSTANT */
    bw_put(w, 0, 8);
    bw_put(w, (uint64_t)(x[0] & 0xFFFF), 16);
    return;
  }
  /* FIXED predictor: the order (0..4) with the smallest sum |residual|, the lowest on ties */
  int best_order = 0;
  uint64_t best_cost = ~0ull;
  memcpy(r, x, (size_t)n * sizeof(int64_t));
  int len = n;
  for (int order = 0; order <= 4; ++order) {
    if (order) {
      for (int i = 0; i + 1 < len; ++i) r[i] = r[i + 1] - r[i];
      --len;
    }
    if (n <= order) break;
    uint64_t cost = 0;
    for (int i = 0; i < len; ++i) cost += (uint64_t)(r[i] < 0 ? -r[i] : r[i]);
    if (cost < best_cost) { best_cost = cost; best_order = order; }
  }
  memcpy(r, x, (size_t)n * sizeof(int64_t));
  len = n;
  for (int order = 1; order <= best_order; ++order) {
    for (int i = 0; i + 1 < len; ++i) r[i] = r[i + 1] - r[i];
    --len;
  }
  uint64_t sum = 0, qmax = 0;
  for (int i = 0; i < len; ++i) {
    const int64_t v = r[i];
    r[i] = v >= 0 ? (v << 1) : (((-v) << 1) - 1);                 /* zig-zag */
    sum += (uint64_t)r[i];
  }
  const double mean = len ? (double)sum / (double)len : 0.0;
  int k = 0;
  while (k < 14 && (double)(1 << (k + 1)) < mean + 1.0) ++k;
  for (int i = 0; i < len; ++i) { const uint64_t q = (uint64_t)r[i] >> k; if (q > qmax) qmax = q; }
  if (len && qmax > 4096) {                                       /* pathological block: VERBATIM */
    bw_put(w, 1u << 1, 8);
    for (int i = 0; i < n; ++i) bw_put(w, (uint64_t)(x[i] & 0xFFFF), 16);
    return;
  }
  bw_put(w, (uint64_t)((8 + best_order) << 1), 8);
  for (int i = 0; i < best_order; ++i) bw_put(w, (uint64_t)(x[i] & 0xFFFF), 16);
  bw_put(w, 0, 2);                                                /* Rice, 4-bit parameters */
  bw_put(w, 0, 4);                                                /* partition order 0 */
  bw_put(w, (uint64_t)k, 4);
  for (int i = 0; i < len; ++i) {
    const uint64_t u = (uint64_t)r[i];
    bw_zeros_then_one(w, u >> k);
    bw_put(w, u & ((1ull << k) - 1), k);
  }
}

int64_t aip_flac_encode16(const int16_t* pcm, int64_t n, int32_t channels, int32_t sample_rate, int32_t blocksize,
                          const uint8_t md5[16], uint8_t* out, size_t cap) {
  if (!pcm || !out || !md5 || n < 0 || channels < 1 || channels > 8 || blocksize < 16 || blocksize > 65535 ||
      sample_rate <= 0 || sample_rate >= (1 << 20))
    return AIP_CODEC_ERR_ARG;
  if (cap < 42) return AIP_CODEC_ERR_CAPACITY;
  int sr_code = 0;
  switch (sample_rate) {
    case 8000: sr_code = 4; break;   case 16000: sr_code = 5; break;  case 22050: sr_code = 6; break;
    case 24000: sr_code = 7; break;  case 32000: sr_code = 8; break;  case 44100: sr_code = 9; break;
    case 48000: sr_code = 10; break; case 96000: sr_code = 11; break; default: sr_code = 0;
  }
  int64_t* x = (int64_t*)malloc((size_t)blocksize * 2 * sizeof(int64_t));
  if (!x) return AIP_CODEC_ERR_ARG;
  int64_t* r = x + blocksize;
  size_t pos = 42;                                                /* "fLaC" + block header + STREAMINFO */
  size_t min_f = (size_t)1 << 24, max_f = 0;
  int64_t frame_no = 0;
  int rc = 0;
  for (int64_t s = 0; s < n; s += blocksize, ++frame_no) {
    const int bs = (int)(n - s < blocksize ? n - s : blocksize);
    const size_t f0 = pos;
    /* header */
    int bs_code = 7;
    for (int c = 1; c < 16; ++c) if (kBlockSizes[c] == bs) bs_code = c;
    uint8_t hdr[16];
    int h = 0;
    hdr[h++] = 0xFF; hdr[h++] = 0xF8;
    hdr[h++] = (uint8_t)((bs_code << 4) | sr_code);
    hdr[h++] = (uint8_t)(((channels - 1) << 4) | (4 << 1));
    {                                                             /* frame number, UTF-8 style */
      uint64_t v = (uint64_t)frame_no;
      if (v < 0x80) hdr[h++] = (uint8_t)v;
      else {
        uint8_t tail[8];
        int nt = 0;
        for (;;) {
          tail[nt++] = (uint8_t)(0x80 | (v & 0x3F));
          v >>= 6;
          if (v < (uint64_t)(0x40 >> nt)) break;
        }
        hdr[h++] = (uint8_t)(((0xFF << (7 - nt)) & 0xFF) | v);
        for (int i = nt - 1; i >= 0; --i) hdr[h++] = tail[i];
      }
    }
    if (bs_code == 7) { hdr[h++] = (uint8_t)((bs - 1) >> 8); hdr[h++] = (uint8_t)((bs - 1) & 0xFF); }
    uint8_t c8 = 0;
    for (int i = 0; i < h; ++i) c8 = crc8_tab[c8 ^ hdr[i]];
    hdr[h++] = c8;
    if (pos + (size_t)h > cap) { rc = AIP_CODEC_ERR_CAPACITY; break; }
    memcpy(out + pos, hdr, (size_t)h);
    pos += (size_t)h;
    bitw w = {out + pos, cap - pos, 0, 0, 0, 0};
    for (int c = 0; c < channels; ++c) encode_subframe(&w, pcm + s * channels + c, bs, channels, x, r);
    bw_flush(&w);
    if (w.err) { rc = AIP_CODEC_ERR_CAPACITY; break; }
    pos += w.pos;
    if (pos + 2 > cap) { rc = AIP_CODEC_ERR_CAPACITY; break; }
    uint16_t c16 = 0;
    for (size_t i = f0; i < pos; ++i) c16 = (uint16_t)((c16 << 8) ^ crc16_tab[(c16 >> 8) ^ out[i]]);
    out[pos++] = (uint8_t)(c16 >> 8);
    out[pos++] = (uint8_t)(c16 & 0xFF);
    const size_t flen = pos - f0;
    if (flen < min_f) min_f = flen;
    if (flen > max_f) max_f = flen;
  }
  free(x);
  if (rc) return rc;
  if (frame_no == 0) min_f = max_f = 0;
  /* stream marker + STREAMINFO (last metadata block) */
  memcpy(out, "fLaC", 4);
  out[4] = 0x80; out[5] = 0; out[6] = 0; out[7] = 34;
  uint8_t* si = out + 8;
  si[0] = (uint8_t)(blocksize >> 8); si[1] = (uint8_t)blocksize; si[2] = si[0]; si[3] = si[1];
  si[4] = (uint8_t)(min_f >> 16); si[5] = (uint8_t)(min_f >> 8); si[6] = (uint8_t)min_f;
  si[7] = (uint8_t)(max_f >> 16); si[8] = (uint8_t)(max_f >> 8); si[9] = (uint8_t)max_f;
  const uint64_t xw = ((uint64_t)sample_rate << 44) | ((uint64_t)(channels - 1) << 41) | ((uint64_t)15 << 36) | (uint64_t)n;
  for (int i = 0; i < 8; ++i) si[10 + i] = (uint8_t)(xw >> (56 - 8 * i));
  memcpy(si + 18, md5, 16);
  return (int64_t)pos;
}

const char* aip_codec_status_string(int status) {
  switch (status) {
    case 0: return "ok";
    case AIP_CODEC_ERR_ARG: return "invalid argument";
    case AIP_CODEC_ERR_FORMAT: return "not a FLAC stream / reserved or inconsistent field";
    case AIP_CODEC_ERR_TRUNCATED: return "stream ends inside a frame";
    case AIP_CODEC_ERR_SYNC: return "lost frame sync";
    case AIP_CODEC_ERR_CAPACITY: return "output buffer too small";
    default: return "unknown status";
  }
}
