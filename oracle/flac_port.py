"""Pure-Python FLAC codec: the checker of the native host codec (csrc/aip_flac.c), TEST INFRASTRUCTURE like the rest of oracle/.

It was the product codec until the native one replaced it (0.4 - 0.9 s per LibriSpeech file against a few milliseconds); the
native decoder must reproduce its PCM sample for sample and the native encoder its streams byte for byte (tests/test_codec.py).
It follows the FLAC format specification (RFC 9639), not a reference source file: the reference reaches its codec through
soundfile / libsndfile (utils.py:36, :87), which is not vendored.

  * decode: fixed / variable block size, CONSTANT / VERBATIM / FIXED / LPC subframes, Rice + escape partitions, left/side,
    right/side, mid/side, self-verified against the STREAMINFO MD5 on request;
  * encode: 16-bit, fixed predictors + Rice, one partition per block.
"""
from __future__ import annotations

import hashlib
import struct
from typing import Optional, Tuple

import numpy as np

__all__ = ["decode_flac", "encode_flac", "FlacInfo"]


# ----------------------------------------------------------------------------- FLAC decode
class FlacInfo:
    __slots__ = ("sample_rate", "channels", "bits_per_sample", "total_samples", "md5",
                 "min_blocksize", "max_blocksize")

    def __repr__(self):
        return (f"FlacInfo(sr={self.sample_rate}, ch={self.channels}, bps={self.bits_per_sample}, "
                f"n={self.total_samples}, block={self.min_blocksize}..{self.max_blocksize})")


_BLOCKSIZE_TABLE = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608,
                    8: 256, 9: 512, 10: 1024, 11: 2048, 12: 4096, 13: 8192, 14: 16384, 15: 32768}
_SAMPLE_SIZE_TABLE = {1: 8, 2: 12, 4: 16, 5: 20, 6: 24, 7: 32}
_FIXED_COEFFS = {0: (), 1: (1,), 2: (2, -1), 3: (3, -3, 1), 4: (4, -6, 4, -1)}


class _BitReader:
    def __init__(self, data: bytes, start_byte: int):
        self.data = data
        self.pos = start_byte * 8
        arr = np.frombuffer(data, dtype=np.uint8)
        self.bits = np.unpackbits(arr)
        ones = np.flatnonzero(self.bits)
        # nxt[p] = position of the first set bit at or after p (sentinel: len(bits))
        nxt = np.full(self.bits.shape[0] + 1, self.bits.shape[0], dtype=np.int64)
        if ones.size:
            idx = np.searchsorted(ones, np.arange(self.bits.shape[0]), side="left")
            ok = idx < ones.size
            nxt[:-1][ok] = ones[idx[ok]]
        self.nxt = nxt

    def read(self, n: int) -> int:
        if n == 0:
            return 0
        p = self.pos
        self.pos = p + n
        b0 = p >> 3
        nb = ((p + n + 7) >> 3) - b0
        v = int.from_bytes(self.data[b0:b0 + nb], "big")
        v >>= nb * 8 - (p & 7) - n
        return v & ((1 << n) - 1)

    def read_signed(self, n: int) -> int:
        v = self.read(n)
        return v - (1 << n) if v >> (n - 1) else v

    def read_unary(self) -> int:
        p = self.pos
        e = int(self.nxt[p])
        self.pos = e + 1
        return e - p

    def align(self):
        self.pos = (self.pos + 7) & ~7

    def rice_block(self, n: int, k: int) -> np.ndarray:
        """n Rice(k)-coded signed residuals starting at the current position."""
        if n <= 0:
            return np.zeros(0, dtype=np.int64)
        nxt = self.nxt
        p = self.pos
        step = k + 1
        ends = np.empty(n, dtype=np.int64)
        for i in range(n):
            e = nxt[p]
            ends[i] = e
            p = e + step
        p0 = self.pos
        self.pos = int(p)
        starts = np.empty(n, dtype=np.int64)
        starts[0] = p0
        starts[1:] = ends[:-1] + step
        u = ends - starts
        if k:
            idx = (ends + 1)[:, None] + np.arange(k)[None, :]
            weights = (1 << np.arange(k - 1, -1, -1)).astype(np.int64)
            rem = self.bits[idx].astype(np.int64) @ weights
            u = (u << k) | rem
        return (u >> 1) ^ -(u & 1)


def _read_utf8_number(br: _BitReader) -> int:
    b = br.read(8)
    if b < 0x80:
        return b
    n = 0
    while b & (0x80 >> n):
        n += 1
    v = b & ((1 << (7 - n)) - 1)
    for _ in range(n - 1):
        v = (v << 6) | (br.read(8) & 0x3F)
    return v


def _residual(br: _BitReader, blocksize: int, order: int) -> np.ndarray:
    method = br.read(2)
    if method > 1:
        raise ValueError("FLAC: reserved residual coding method")
    pbits = 4 if method == 0 else 5
    escape = (1 << pbits) - 1
    porder = br.read(4)
    nparts = 1 << porder
    out = np.empty(blocksize - order, dtype=np.int64)
    w = 0
    for part in range(nparts):
        n = (blocksize >> porder) - (order if part == 0 else 0)
        k = br.read(pbits)
        if k == escape:
            raw = br.read(5)
            vals = np.fromiter((br.read_signed(raw) if raw else 0 for _ in range(n)), dtype=np.int64, count=n)
        else:
            vals = br.rice_block(n, k)
        out[w:w + n] = vals
        w += n
    return out


def _subframe(br: _BitReader, blocksize: int, bps: int) -> np.ndarray:
    if br.read(1):
        raise ValueError("FLAC: subframe padding bit set")
    typ = br.read(6)
    wasted = 0
    if br.read(1):
        wasted = br.read_unary() + 1
        bps -= wasted
    if typ == 0:                                    # CONSTANT
        out = np.full(blocksize, br.read_signed(bps), dtype=np.int64)
    elif typ == 1:                                  # VERBATIM
        out = np.fromiter((br.read_signed(bps) for _ in range(blocksize)), dtype=np.int64, count=blocksize)
    elif 8 <= typ <= 12:                            # FIXED
        order = typ - 8
        warm = [br.read_signed(bps) for _ in range(order)]
        res = _residual(br, blocksize, order)
        out = _restore_lpc(warm, res, _FIXED_COEFFS[order], 0, blocksize)
    elif typ >= 32:                                 # LPC
        order = (typ & 31) + 1
        warm = [br.read_signed(bps) for _ in range(order)]
        precision = br.read(4) + 1
        if precision == 16:
            raise ValueError("FLAC: invalid LPC precision")
        shift = br.read_signed(5)
        coefs = [br.read_signed(precision) for _ in range(order)]
        res = _residual(br, blocksize, order)
        out = _restore_lpc(warm, res, coefs, shift, blocksize)
    else:
        raise ValueError(f"FLAC: reserved subframe type {typ}")
    if wasted:
        out = out << wasted
    return out


def _restore_lpc(warm, res: np.ndarray, coefs, shift: int, blocksize: int) -> np.ndarray:
    order = len(coefs)
    if order == 0:
        return res.copy()
    s = list(warm) + [0] * (blocksize - order)
    r = res.tolist()
    if order == 1 and shift == 0 and coefs[0] == 1:
        out = np.empty(blocksize, dtype=np.int64)
        out[0] = warm[0]
        out[1:] = warm[0] + np.cumsum(res)
        return out
    c = list(coefs)
    # s[n] = r[n-order] + (sum_i c[i] * s[n-1-i]) >> shift    (arithmetic shift == floor)
    if order == 2:
        c0, c1 = c
        a, b = s[1], s[0]
        for n in range(2, blocksize):
            v = r[n - 2] + ((c0 * a + c1 * b) >> shift)
            s[n] = v
            b = a
            a = v
    else:
        rng_o = range(order)
        for n in range(order, blocksize):
            acc = 0
            for i in rng_o:
                acc += c[i] * s[n - 1 - i]
            s[n] = r[n - order] + (acc >> shift)
    return np.asarray(s, dtype=np.int64)


def decode_flac(data: bytes, max_samples: Optional[int] = None, verify_md5: bool = False
                ) -> Tuple[np.ndarray, FlacInfo]:
    """Decode a FLAC byte string to int PCM [n, channels] (int32) plus stream info.

    ``max_samples`` stops after the frame that reaches that many samples (load_audio only
    needs the first ``sr * max_len``).  ``verify_md5`` decodes everything and compares
    with the STREAMINFO signature, raising ValueError on mismatch.
    """
    if data[:4] != b"fLaC":
        raise ValueError("not a FLAC stream")
    pos = 4
    info = None
    while True:
        hdr = data[pos]
        last, btype = hdr >> 7, hdr & 0x7F
        length = int.from_bytes(data[pos + 1:pos + 4], "big")
        body = data[pos + 4:pos + 4 + length]
        pos += 4 + length
        if btype == 0:
            info = FlacInfo()
            info.min_blocksize, info.max_blocksize = struct.unpack(">HH", body[:4])
            x = int.from_bytes(body[10:18], "big")
            info.sample_rate = x >> 44
            info.channels = ((x >> 41) & 7) + 1
            info.bits_per_sample = ((x >> 36) & 31) + 1
            info.total_samples = x & ((1 << 36) - 1)
            info.md5 = body[18:34]
        if last:
            break
    if info is None:
        raise ValueError("FLAC: missing STREAMINFO")
    if verify_md5:
        max_samples = None
    br = _BitReader(data, pos)
    nbits = len(data) * 8
    chunks = []
    got = 0
    want = info.total_samples if info.total_samples else None
    while br.pos + 16 <= nbits and (want is None or got < want):
        if max_samples is not None and got >= max_samples:
            break
        sync = br.read(14)
        if sync != 0x3FFE:
            raise ValueError("FLAC: lost frame sync")
        br.read(1)
        br.read(1)                                   # blocking strategy (number parsed either way)
        bs_code = br.read(4)
        sr_code = br.read(4)
        ch_code = br.read(4)
        ss_code = br.read(3)
        br.read(1)
        _read_utf8_number(br)
        if bs_code == 6:
            blocksize = br.read(8) + 1
        elif bs_code == 7:
            blocksize = br.read(16) + 1
        elif bs_code in _BLOCKSIZE_TABLE:
            blocksize = _BLOCKSIZE_TABLE[bs_code]
        else:
            raise ValueError("FLAC: reserved block size code")
        if sr_code == 12:
            br.read(8)
        elif sr_code in (13, 14):
            br.read(16)
        br.read(8)                                   # CRC-8 (the MD5 check covers integrity)
        bps = _SAMPLE_SIZE_TABLE.get(ss_code, info.bits_per_sample) if ss_code else info.bits_per_sample
        if ch_code < 8:
            nch = ch_code + 1
            subs = [_subframe(br, blocksize, bps) for _ in range(nch)]
        elif ch_code == 8:                           # left / side
            left = _subframe(br, blocksize, bps)
            side = _subframe(br, blocksize, bps + 1)
            subs = [left, left - side]
        elif ch_code == 9:                           # side / right
            side = _subframe(br, blocksize, bps + 1)
            right = _subframe(br, blocksize, bps)
            subs = [right + side, right]
        elif ch_code == 10:                          # mid / side
            mid = _subframe(br, blocksize, bps)
            side = _subframe(br, blocksize, bps + 1)
            mid = (mid << 1) | (side & 1)
            subs = [(mid + side) >> 1, (mid - side) >> 1]
        else:
            raise ValueError("FLAC: reserved channel assignment")
        br.align()
        br.read(16)                                  # CRC-16
        chunks.append(np.stack(subs, axis=1))
        got += blocksize
    pcm = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, info.channels), dtype=np.int64)
    if want is not None:
        pcm = pcm[:want]
    if verify_md5 and any(info.md5):
        nbytes = (info.bits_per_sample + 7) // 8
        if nbytes == 2:
            raw = pcm.astype("<i2").tobytes()
        elif nbytes == 1:
            raw = pcm.astype("i1").tobytes()
        elif nbytes == 4:
            raw = pcm.astype("<i4").tobytes()
        else:                                        # 24-bit: three little-endian bytes per sample
            raw = pcm.astype("<i4").reshape(-1).view(np.uint8).reshape(-1, 4)[:, :3].tobytes()
        if hashlib.md5(raw).digest() != info.md5:
            raise ValueError("FLAC: MD5 signature mismatch")
    return pcm.astype(np.int32), info


# ----------------------------------------------------------------------------- FLAC encode
def _crc_table(poly: int, bits: int):
    top = 1 << (bits - 1)
    mask = (1 << bits) - 1
    tab = []
    for i in range(256):
        c = i << (bits - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
        tab.append(c)
    return tab


_CRC8 = _crc_table(0x07, 8)
_CRC16 = _crc_table(0x8005, 16)


def _crc8(b: bytes) -> int:
    c = 0
    for x in b:
        c = _CRC8[c ^ x]
    return c


def _crc16(b: bytes) -> int:
    c = 0
    for x in b:
        c = ((c << 8) & 0xFFFF) ^ _CRC16[(c >> 8) ^ x]
    return c


def _utf8_number(v: int) -> bytes:
    if v < 0x80:
        return bytes([v])
    out = []
    n = 0
    while True:
        out.append(0x80 | (v & 0x3F))
        v >>= 6
        n += 1
        if v < (0x40 >> n):
            break
    lead = ((0xFF << (7 - n)) & 0xFF) | v
    return bytes([lead] + out[::-1])


def _bits_of(values: np.ndarray, width: int) -> np.ndarray:
    """Big-endian ``width``-bit two's-complement fields of ``values`` as a flat 0/1 array."""
    v = values.astype(np.int64) & ((1 << width) - 1)
    shifts = np.arange(width - 1, -1, -1, dtype=np.int64)
    return ((v[:, None] >> shifts[None, :]) & 1).astype(np.uint8).reshape(-1)


def _encode_subframe(x: np.ndarray, bps: int) -> np.ndarray:
    """One FIXED-predictor subframe (best order 0..4, Rice partition order 0) as a bit array."""
    n = x.shape[0]
    x = x.astype(np.int64)
    if n and np.all(x == x[0]):
        return np.concatenate([_bits_of(np.array([0]), 8), _bits_of(x[:1], bps)])     # CONSTANT
    best = None
    res = x
    for order in range(0, 5):
        if order:
            res = np.diff(res)
        if n <= order:
            break
        cost = int(np.abs(res).sum())
        if best is None or cost < best[0]:
            best = (cost, order, res.copy())
    _, order, res = best
    u = np.where(res >= 0, res << 1, ((-res) << 1) - 1)          # zig-zag
    mean = float(u.mean()) if u.size else 0.0
    k = 0
    while k < 14 and (1 << (k + 1)) < mean + 1:
        k += 1
    q = u >> k
    if q.size and int(q.max()) > 4096:                           # pathological: store verbatim
        return np.concatenate([_bits_of(np.array([1 << 1]), 8), _bits_of(x, bps)])
    head = [_bits_of(np.array([(8 + order) << 1]), 8), _bits_of(x[:order], bps),
            _bits_of(np.array([0]), 2), _bits_of(np.array([0]), 4), _bits_of(np.array([k]), 4)]
    lens = q + 1 + k
    ends = np.cumsum(lens)
    starts = ends - lens
    body = np.zeros(int(ends[-1]) if ends.size else 0, dtype=np.uint8)
    body[starts + q] = 1
    for j in range(k):
        body[starts + q + 1 + j] = (u >> (k - 1 - j)) & 1
    return np.concatenate(head + [body])


def encode_flac(pcm: np.ndarray, sample_rate: int, bits_per_sample: int = 16, blocksize: int = 4096) -> bytes:
    """Encode int PCM [n] or [n, channels] as a FLAC stream (independent channels)."""
    pcm = np.asarray(pcm)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    n, nch = pcm.shape
    if bits_per_sample != 16:
        raise ValueError("encode_flac writes 16-bit streams only")
    sr_code = {8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}.get(sample_rate, 0)
    frames = []
    min_f, max_f = 1 << 24, 0
    for fi, s in enumerate(range(0, n, blocksize)):
        blk = pcm[s:s + blocksize]
        bs = blk.shape[0]
        if bs in (192, 576, 1152, 2304, 4608, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
            bs_code = {v: k for k, v in _BLOCKSIZE_TABLE.items()}[bs]
            extra = b""
        else:
            bs_code, extra = 7, struct.pack(">H", bs - 1)
        hdr = bytes([0xFF, 0xF8, (bs_code << 4) | sr_code, ((nch - 1) << 4) | (4 << 1)])
        hdr += _utf8_number(fi) + extra
        hdr += bytes([_crc8(hdr)])
        bits = np.concatenate([_encode_subframe(blk[:, c], bits_per_sample) for c in range(nch)])
        pad = (-bits.shape[0]) % 8
        if pad:
            bits = np.concatenate([bits, np.zeros(pad, dtype=np.uint8)])
        frame = hdr + np.packbits(bits).tobytes()
        frame += struct.pack(">H", _crc16(frame))
        frames.append(frame)
        min_f, max_f = min(min_f, len(frame)), max(max_f, len(frame))
    if not frames:
        min_f = max_f = 0
    md5 = hashlib.md5(pcm.astype("<i2").tobytes()).digest()
    x = (sample_rate << 44) | ((nch - 1) << 41) | ((bits_per_sample - 1) << 36) | n
    streaminfo = struct.pack(">HH", blocksize, blocksize) + min_f.to_bytes(3, "big") + max_f.to_bytes(3, "big")
    streaminfo += x.to_bytes(8, "big") + md5
    return b"fLaC" + bytes([0x80]) + len(streaminfo).to_bytes(3, "big") + streaminfo + b"".join(frames)


