"""CPU tier: the native host codec (csrc/aip_flac.c, include/aip_codec.h) against the pure-Python codec in oracle/flac_port.py,
on streams written by a small test-side FLAC writer that exercises every decoder path (LPC, wasted bits, the three stereo
decorrelations, escape partitions, partition orders, 8 / 16 / 24-bit, odd block sizes), on the reference's own files when this
container has them, and byte for byte on the encoder."""
import hashlib
import re
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import pytest

from ml_audio_inpainting_b200 import _codec, audio_io
from oracle import flac_port

ROOT = Path(__file__).resolve().parents[1]
REF_SAMPLES = Path("/root/reference/test_samples")


# ------------------------------------------------------------------------------------------- a test-side FLAC writer
class Bits:
    def __init__(self):
        self.b = []

    def put(self, v, n):
        v &= (1 << n) - 1
        self.b.extend((v >> i) & 1 for i in range(n - 1, -1, -1))

    def unary(self, z):
        self.b.extend([0] * z + [1])

    def pad(self):
        self.b.extend([0] * (-len(self.b) % 8))

    def bytes(self):
        assert len(self.b) % 8 == 0
        return np.packbits(np.array(self.b, dtype=np.uint8)).tobytes()


def put_residual(w, res, blocksize, order, porder, ks, method=0):
    """ks[p] = Rice parameter of partition p, or ('esc', raw_bits)."""
    pbits = 4 if method == 0 else 5
    w.put(method, 2)
    w.put(porder, 4)
    i = 0
    for p in range(1 << porder):
        n = (blocksize >> porder) - (order if p == 0 else 0)
        k = ks[p]
        if isinstance(k, tuple):
            w.put((1 << pbits) - 1, pbits)
            w.put(k[1], 5)
            for v in res[i:i + n]:
                if k[1]:
                    w.put(int(v), k[1])
        else:
            w.put(k, pbits)
            for v in res[i:i + n]:
                u = (int(v) << 1) if v >= 0 else ((-int(v)) << 1) - 1
                w.unary(u >> k)
                if k:
                    w.put(u, k)
        i += n
    assert i == len(res)


def put_subframe(w, x, bps, spec):
    """spec: ('const',) | ('verbatim',) | ('fixed', order, porder, ks) | ('lpc', coefs, precision, shift, porder, ks[, method]);
    optional trailing {'wasted': k}."""
    opts = spec[-1] if isinstance(spec[-1], dict) else {}
    wasted = opts.get("wasted", 0)
    x = [int(v) for v in x]
    if wasted:
        assert all(v % (1 << wasted) == 0 for v in x)
        x = [v >> wasted for v in x]
        bps -= wasted
    kind = spec[0]
    typ = {"const": 0, "verbatim": 1}.get(kind)
    if kind == "fixed":
        typ = 8 + spec[1]
    if kind == "lpc":
        typ = 32 + len(spec[1]) - 1
    w.put(0, 1)
    w.put(typ, 6)
    w.put(1 if wasted else 0, 1)
    if wasted:
        w.unary(wasted - 1)
    n = len(x)
    if kind == "const":
        w.put(x[0], bps)
    elif kind == "verbatim":
        for v in x:
            w.put(v, bps)
    else:
        if kind == "fixed":
            order = spec[1]
            coefs, shift = list(flac_port._FIXED_COEFFS[order]), 0
            porder, ks, method = spec[2], spec[3], 0
        else:
            coefs, precision, shift, porder, ks = spec[1], spec[2], spec[3], spec[4], spec[5]
            method = spec[6] if len(spec) > 6 and not isinstance(spec[6], dict) else 0
            order = len(coefs)
        for v in x[:order]:
            w.put(v, bps)
        if kind == "lpc":
            w.put(precision - 1, 4)
            w.put(shift, 5)
            for c in coefs:
                w.put(c, precision)
        res = [x[i] - (sum(c * x[i - 1 - j] for j, c in enumerate(coefs)) >> shift) for i in range(order, n)]
        put_residual(w, res, n, order, porder, ks, method)


def make_stream(pcm, bps, sr, frames):
    """pcm [n, ch] ints; frames: list of (blocksize, ch_mode, [subframe spec per coded channel])."""
    pcm = np.asarray(pcm, dtype=np.int64)
    n, nch = pcm.shape
    body = b""
    s = 0
    ss_code = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6}[bps]
    for fi, (bs, mode, specs) in enumerate(frames):
        blk = pcm[s:s + bs]
        assert len(blk) == bs
        hdr = Bits()
        hdr.put(0x3FFE, 14); hdr.put(0, 1); hdr.put(0, 1)
        table = {v: k for k, v in flac_port._BLOCKSIZE_TABLE.items()}
        bs_code = table.get(bs, 6 if bs <= 256 else 7)
        hdr.put(bs_code, 4); hdr.put(0, 4)
        hdr.put({"indep": nch - 1, "left_side": 8, "side_right": 9, "mid_side": 10}[mode], 4)
        hdr.put(ss_code, 3); hdr.put(0, 1)
        hb = hdr.bytes() + flac_port._utf8_number(fi)
        if bs_code == 6:
            hb += bytes([bs - 1])
        elif bs_code == 7:
            hb += (bs - 1).to_bytes(2, "big")
        hb += bytes([flac_port._crc8(hb)])
        w = Bits()
        if mode == "indep":
            chans = [(blk[:, c], bps) for c in range(nch)]
        else:
            l, r = blk[:, 0], blk[:, 1]
            side = l - r
            if mode == "left_side":
                chans = [(l, bps), (side, bps + 1)]
            elif mode == "side_right":
                chans = [(side, bps + 1), (r, bps)]
            else:
                chans = [((l + r) >> 1, bps), (side, bps + 1)]
        for (x, b), spec in zip(chans, specs):
            put_subframe(w, x, b, spec)
        w.pad()
        fr = hb + w.bytes()
        fr += flac_port._crc16(fr).to_bytes(2, "big")
        body += fr
        s += bs
    assert s == n
    nbytes = (bps + 7) // 8
    raw = b"".join(int(v).to_bytes(nbytes, "little", signed=True) for v in pcm.reshape(-1))
    x = (sr << 44) | ((nch - 1) << 41) | ((bps - 1) << 36) | n
    bmax = max(f[0] for f in frames)
    si = (min(f[0] for f in frames)).to_bytes(2, "big") + bmax.to_bytes(2, "big") + bytes(6) + x.to_bytes(8, "big")
    si += hashlib.md5(raw).digest()
    return b"fLaC" + bytes([0x80]) + len(si).to_bytes(3, "big") + si + body


def speechlike(n, nch, bps, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    amp = (1 << (bps - 1)) * 0.3
    x = np.stack([amp * (np.sin(0.03 * t + c) + 0.3 * np.sin(0.41 * t)) + amp * 0.02 * rng.standard_normal(n) for c in range(nch)], 1)
    return np.rint(x).astype(np.int64)


def both(stream, **kw):
    got, gi = audio_io.decode_flac(stream, **kw)
    ref, ri = flac_port.decode_flac(stream, **kw)
    assert np.array_equal(got, ref)
    assert (gi.sample_rate, gi.channels, gi.bits_per_sample, gi.total_samples, gi.md5, gi.min_blocksize, gi.max_blocksize) == \
           (ri.sample_rate, ri.channels, ri.bits_per_sample, ri.total_samples, ri.md5, ri.min_blocksize, ri.max_blocksize)
    return got


# ------------------------------------------------------------------------------------------- tests
def test_header_and_binding_agree():
    text = (ROOT / "include" / "aip_codec.h").read_text()
    declared = set(re.findall(r"\b(aip_[a-z0-9_]+)\s*\(", text))
    assert declared == set(_codec.SIGNATURES), declared ^ set(_codec.SIGNATURES)
    lib = _codec.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.aip_codec_status_string(0) == b"ok" and b"sync" in lib.aip_codec_status_string(-4)


def test_decoder_paths_against_the_python_codec():
    lpc8 = [1412, -1203, 601, -322, 188, -97, 41, -12]                   # precision 12, shift 10: any coefficients are a valid stream
    x = speechlike(4096 + 1152 + 200 + 16, 1, 16, 1)
    frames = [(4096, "indep", [("lpc", lpc8, 12, 10, 3, [4, 5, ("esc", 14), 3, 6, 4, 5, 7])]),
              (1152, "indep", [("fixed", 2, 1, [5, 6])]),
              (200, "indep", [("lpc", [3, -3, 1], 5, 0, 0, [9], 1)]),    # 5-bit Rice parameters (method 1), variable block size
              (16, "indep", [("verbatim",)])]
    s1 = make_stream(x, 16, 16000, frames)
    assert np.array_equal(both(s1, verify_md5=True), x)
    assert both(s1, max_samples=4096).shape[0] == 4096 and both(s1, max_samples=4097).shape[0] == 4096 + 1152
    # stereo, all three decorrelations, wasted bits, a constant channel, 24-bit
    y = speechlike(3 * 576, 2, 24, 2)
    y[:576, 1] = (y[:576, 1] >> 3) << 3                                  # three wasted bits in the right channel of frame 0
    y[2 * 576:, 0] = 1234
    frames = [(576, "indep", [("fixed", 4, 2, [13, 12, 13, 12]), ("fixed", 1, 0, [11], {"wasted": 3})]),
              (576, "mid_side", [("lpc", [1900, -900], 12, 10, 0, [12]), ("fixed", 2, 0, [10])]),
              (576, "indep", [("const",), ("fixed", 0, 0, [14])])]
    s2 = make_stream(y, 24, 48000, frames)
    assert np.array_equal(both(s2, verify_md5=True), y)
    z = speechlike(2 * 192, 2, 8, 3)
    frames = [(192, "left_side", [("fixed", 1, 0, [3]), ("fixed", 3, 1, [4, ("esc", 9)])]),
              (192, "side_right", [("fixed", 0, 0, [6]), ("verbatim",)])]
    s3 = make_stream(z, 8, 8000, frames)
    assert np.array_equal(both(s3, verify_md5=True), z)


def test_the_references_own_files_when_present(golden_clips):
    if not REF_SAMPLES.is_dir():
        pytest.skip("/root/reference is not on this box")
    files = sorted(REF_SAMPLES.glob("*.flac"))
    assert len(files) == 9
    z = np.load(ROOT / "tests" / "golden" / "clips_int16.npz")
    for f in files:
        data = f.read_bytes()
        pcm, info = audio_io.decode_flac(data, verify_md5=True)              # libFLAC-written: LPC subframes, Rice partitions
        assert info.sample_rate == 16000 and info.channels == 1 and pcm.shape[0] == info.total_samples
        assert np.array_equal(pcm[:80000, 0], z[f.stem].astype(np.int32))
        head = both(data, max_samples=80000)
        assert 80000 <= head.shape[0] < 80000 + 4096 and np.array_equal(head, pcm[:head.shape[0]])


def test_encoder_is_byte_identical_and_round_trips(golden_clips):
    rng = np.random.default_rng(0)
    cases = [np.rint(golden_clips[k][:20000] * 32768).astype(np.int16) for k in sorted(golden_clips)[:3]]
    cases += [np.zeros(5000, np.int16), np.full(4096, -7, np.int16), rng.integers(-32768, 32768, 9000).astype(np.int16),
              np.array([5], np.int16), np.array([1, -2, 3], np.int16), np.zeros(0, np.int16),
              np.stack([cases[0][:6000], cases[1][:6000]], 1),
              (3000 * np.sin(np.arange(10000) * 0.01)).astype(np.int16)]
    for x in cases:
        for bs in (4096, 1152, 1000):
            blob = audio_io.encode_flac(x, 16000, blocksize=bs)
            assert blob == flac_port.encode_flac(x, 16000, blocksize=bs), (x.shape, bs)
            back, info = audio_io.decode_flac(blob, verify_md5=True)
            assert np.array_equal(back, x.reshape(len(x), 1 if x.ndim == 1 else x.shape[1])) and info.total_samples == len(x)


def test_bad_streams_raise():
    x = (1000 * np.sin(np.arange(9000) * 0.05)).astype(np.int16)
    blob = audio_io.encode_flac(x, 16000)
    with pytest.raises(ValueError):
        audio_io.decode_flac(b"RIFF" + blob[4:])
    with pytest.raises(ValueError):
        audio_io.decode_flac(blob[:len(blob) // 2])                          # ends inside a frame
    bad = bytearray(blob)
    bad[42] ^= 0xFF                                                          # first frame's sync code
    with pytest.raises(ValueError):
        audio_io.decode_flac(bytes(bad))
    flipped = bytearray(blob)
    flipped[len(blob) // 2] ^= 0x10                                          # a residual bit: decodes, but not to the signed PCM
    try:
        audio_io.decode_flac(bytes(flipped), verify_md5=True)
        raised = False
    except ValueError:
        raised = True
    assert raised


def test_parallel_decodes_agree(golden_clips):
    blobs = [audio_io.encode_flac(np.rint(golden_clips[k] * 32768).astype(np.int16), 16000) for k in sorted(golden_clips)]
    serial = [audio_io.decode_flac(b)[0] for b in blobs]
    with ThreadPoolExecutor(8) as pool:
        par = list(pool.map(lambda b: audio_io.decode_flac(b)[0], blobs * 4))
    for i, p in enumerate(par):
        assert np.array_equal(p, serial[i % len(blobs)])


def test_corrupted_streams_never_crash_the_decoder(golden_clips):
    """Memory safety of the native decoder: random byte corruption, truncation and header damage either decode to something or
    raise ValueError -- never touch memory outside the buffers (a crash would take the interpreter down with it)."""
    rng = np.random.default_rng(11)
    x = np.rint(golden_clips[sorted(golden_clips)[0]][:12000] * 32768).astype(np.int16)
    streams = [audio_io.encode_flac(x, 16000, blocksize=1152), audio_io.encode_flac(np.stack([x, x[::-1]], 1), 16000)]
    lpc8 = [1412, -1203, 601, -322, 188, -97, 41, -12]
    streams.append(make_stream(speechlike(4096, 1, 16, 5), 16, 16000,
                               [(4096, "indep", [("lpc", lpc8, 12, 10, 3, [4, 5, ("esc", 14), 3, 6, 4, 5, 7])])]))
    outcomes = {"ok": 0, "raised": 0}
    for s in streams:
        for trial in range(250):
            b = bytearray(s)
            kind = trial % 3
            if kind == 0:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(4, len(b)))] = int(rng.integers(0, 256))
            elif kind == 1:
                b = b[:int(rng.integers(8, len(b)))]
            else:
                lo = int(rng.integers(42, len(b) - 8))
                b[lo:lo + 8] = bytes(rng.integers(0, 256, 8, dtype=np.uint8))
            try:
                audio_io.decode_flac(bytes(b))
                outcomes["ok"] += 1
            except ValueError:
                outcomes["raised"] += 1
    assert outcomes["raised"] > 100 and sum(outcomes.values()) == 750
