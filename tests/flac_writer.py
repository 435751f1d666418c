"""A FLAC ENCODER for the tests of csrc/flac.cu (test infrastructure; the product only decodes).

Written from the published format (RFC 9639) independently of the decoder — Python integers and numpy here, a C++
bit reader there — so the pair meets only in the byte stream: marker, STREAMINFO (+ PADDING / VORBIS_COMMENT blocks),
frame headers with UTF-8 coded numbers and CRC-8, CONSTANT / VERBATIM / FIXED / LPC subframes, wasted bits,
partitioned Rice residuals (4- and 5-bit parameters, escape partitions), the three stereo decorrelations, CRC-16 and the
MD5 of the samples. No FLAC stream or third-party codec exists in the build image: what pins the decoder beyond this
round trip are the checksums real encoders embed (every CRC is verified by the decoder, the MD5 by audio/reader.py),
which makes any real file self-checking."""
from __future__ import annotations

import hashlib

import numpy as np

RATE_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
SIZE_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}
BLOCK_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13, 16384: 14, 32768: 15}


class BitWriter:
    def __init__(self):
        self.out, self.acc, self.n = bytearray(), 0, 0

    def put(self, value: int, bits: int) -> None:
        if bits == 0:
            return
        self.acc = (self.acc << bits) | (int(value) & ((1 << bits) - 1))
        self.n += bits
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def unary(self, q: int) -> None:
        while q >= 32:
            self.put(0, 32); q -= 32
        self.put(1, q + 1)

    def align(self) -> None:
        if self.n:
            self.put(0, 8 - self.n)


def crc8(data: bytes) -> int:
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def crc16(data: bytes) -> int:
    c = 0
    for b in data:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


def utf8_number(n: int) -> bytes:
    if n < 0x80:
        return bytes([n])
    extra = 1
    while n >= (1 << (5*extra + 6)):
        extra += 1
    lead = ((0xFF << (7 - extra)) & 0xFF) | (n >> (6*extra))
    return bytes([lead] + [0x80 | ((n >> (6*k)) & 0x3F) for k in range(extra - 1, -1, -1)])


def rice_bits(u: np.ndarray, k: int) -> int:
    return int((u >> k).sum()) + u.size*(1 + k)


def put_residual(w: BitWriter, residual: np.ndarray, blocksize: int, order: int, porder: int, rice2: bool, escape: set) -> None:
    w.put(1 if rice2 else 0, 2)
    w.put(porder, 4)
    pbits, top = (5, 30) if rice2 else (4, 14)
    at = 0
    for part in range(1 << porder):
        count = (blocksize >> porder) - (order if part == 0 else 0)
        e = residual[at:at + count].astype(object)
        at += count
        if part in escape:
            width = max([int(v).bit_length() + 1 for v in e] + [0]) if any(int(v) != 0 for v in e) else 0
            w.put((1 << pbits) - 1, pbits)
            w.put(width, 5)
            for v in e:
                w.put(int(v), width)
            continue
        u = np.asarray([(int(v) << 1) ^ (int(v) >> 63) for v in e], dtype=object)
        best = min(range(top + 1), key=lambda k: sum((int(x) >> k) + 1 + k for x in u)) if len(u) else 0
        w.put(best, pbits)
        for x in u:
            w.unary(int(x) >> best)
            w.put(int(x) & ((1 << best) - 1), best)


def lpc_coefficients(x: np.ndarray, order: int, precision: int, shift: int) -> list[int]:
    rows = np.stack([x[order - 1 - j:len(x) - 1 - j] for j in range(order)], axis=1).astype(np.float64)
    target = x[order:].astype(np.float64)
    c = np.linalg.lstsq(rows, target, rcond=None)[0] if len(target) else np.zeros(order)
    limit = (1 << (precision - 1)) - 1
    return [int(np.clip(round(v*(1 << shift)), -limit - 1, limit)) for v in c]


def put_subframe(w: BitWriter, x: np.ndarray, bps: int, kind: str, porder: int, rice2: bool, escape: set, wasted: int) -> None:
    x = [int(v) for v in x]
    n = len(x)
    if wasted:
        assert all(v % (1 << wasted) == 0 for v in x)
        x = [v >> wasted for v in x]
        bps -= wasted
    if kind == "auto":
        kind = "constant" if len(set(x)) == 1 else "fixed2"
    code = {"constant": 0, "verbatim": 1}.get(kind)
    order = 0
    if kind.startswith("fixed"):
        order = int(kind[5:]); code = 8 + order
    elif kind.startswith("lpc"):
        order = int(kind[3:]); code = 32 + order - 1
    w.put(0, 1); w.put(code, 6)
    if wasted:
        w.put(1, 1); w.unary(wasted - 1)
    else:
        w.put(0, 1)
    if kind == "constant":
        w.put(x[0], bps)
    elif kind == "verbatim":
        for v in x:
            w.put(v, bps)
    elif kind.startswith("fixed"):
        for v in x[:order]:
            w.put(v, bps)
        r = np.asarray(x, dtype=object)
        for _ in range(order):
            r = np.concatenate([r[:1]*0, r[1:] - r[:-1]])
        put_residual(w, r[order:], n, order, porder, rice2, escape)
    else:
        precision, shift = 12, 9
        coef = lpc_coefficients(np.asarray(x, dtype=np.float64), order, precision, shift)
        for v in x[:order]:
            w.put(v, bps)
        w.put(precision - 1, 4); w.put(shift, 5)
        for c in coef:
            w.put(c, precision)
        r = [x[i] - (sum(coef[j]*x[i - 1 - j] for j in range(order)) >> shift) for i in range(order, n)]
        put_residual(w, np.asarray(r, dtype=object), n, order, porder, rice2, escape)


def write_flac(samples, rate: int = 44100, bits: int = 16, blocksize: int = 1152, subframe="auto", stereo: str = "independent",
               porder: int = 2, rice2: bool = False, escape=(), wasted: int = 0, id3: bool = False, total_known: bool = True,
               md5: bool = True, size_from_streaminfo: bool = False, first_number: int = 0, trailing: bytes = b"") -> bytes:
    """samples: int array (frames, channels). `subframe`: one kind for every channel or a list per channel
    ('auto' | 'constant' | 'verbatim' | 'fixed0'..'fixed4' | 'lpc1'..'lpc32')"""
    samples = np.asarray(samples).reshape(len(samples), -1)
    frames, channels = samples.shape
    kinds = [subframe]*channels if isinstance(subframe, str) else list(subframe)
    width = (bits + 7)//8
    digest = hashlib.md5(b"".join(int(v).to_bytes(width, "little", signed=True) for v in samples.reshape(-1))).digest() if md5 else bytes(16)
    body = bytearray()
    number = first_number
    for start in range(0, frames, blocksize):
        block = samples[start:start + blocksize]
        n = len(block)
        w = BitWriter()
        w.put(0x3FFE, 14); w.put(0, 1); w.put(0, 1)
        bs_code = BLOCK_CODES.get(n, 6 if n <= 256 else 7)
        sr_code = RATE_CODES.get(rate, 12 if (rate % 1000 == 0 and rate < 256000) else (13 if rate < 65536 else 14))
        assignment = {"independent": channels - 1, "left_side": 8, "side_right": 9, "mid_side": 10}[stereo]
        w.put(bs_code, 4); w.put(sr_code, 4); w.put(assignment, 4); w.put(0 if size_from_streaminfo else SIZE_CODES[bits], 3); w.put(0, 1)
        for b in utf8_number(number):
            w.put(b, 8)
        if bs_code == 6:
            w.put(n - 1, 8)
        elif bs_code == 7:
            w.put(n - 1, 16)
        if sr_code == 12:
            w.put(rate//1000, 8)
        elif sr_code == 13:
            w.put(rate, 16)
        elif sr_code == 14:
            w.put(rate//10, 16)
        w.put(crc8(bytes(w.out)), 8)
        cols = [[int(v) for v in block[:, c]] for c in range(channels)]
        widths = [bits]*channels
        if stereo != "independent":
            left, right = cols
            side = [a - b for a, b in zip(left, right)]
            if stereo == "left_side":
                cols, widths = [left, side], [bits, bits + 1]
            elif stereo == "side_right":
                cols, widths = [side, right], [bits + 1, bits]
            else:
                cols, widths = [[(a + b) >> 1 for a, b in zip(left, right)], side], [bits, bits + 1]
        part = porder
        while part and ((n % (1 << part)) or (n >> part) < 4):
            part -= 1
        for c in range(channels):
            kind = kinds[c]
            if kind != "constant" and kind != "verbatim" and kind != "auto" and n <= int(kind.lstrip("fixedlpc")):
                kind = "verbatim"
            put_subframe(w, cols[c], widths[c], kind, part, rice2, {p for p in escape if p < (1 << part)}, wasted)
        w.align()
        w.put(crc16(bytes(w.out)), 16)
        body += w.out
        number += 1
    info = BitWriter()
    info.put(min(blocksize, 65535), 16); info.put(min(blocksize, 65535), 16); info.put(0, 24); info.put(0, 24)
    info.put(rate, 20); info.put(channels - 1, 3); info.put(bits - 1, 5); info.put(frames if total_known else 0, 36)
    stream = bytearray()
    if id3:
        tag = b"TIT2" + bytes(60)
        stream += b"ID3\x04\x00\x00" + bytes([(len(tag) >> 21) & 0x7F, (len(tag) >> 14) & 0x7F, (len(tag) >> 7) & 0x7F, len(tag) & 0x7F]) + tag
    stream += b"fLaC"
    stream += bytes([0x00, 0, 0, 34]) + bytes(info.out) + digest
    stream += bytes([0x01, 0, 0, 5]) + bytes(5)                                      # PADDING
    comment = (7).to_bytes(4, "little") + b"sfbtest" + (0).to_bytes(4, "little")
    stream += bytes([0x84, 0, 0, len(comment)]) + comment                            # VORBIS_COMMENT, last block
    return bytes(stream + body + trailing)
