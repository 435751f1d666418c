// flac.cu — FLAC stream → interleaved PCM, on the host (audio files without an ffmpeg child; the reference
// reaches every compressed format through `ffmpeg -f f32le`, shaderflow/ffmpeg.py:1240-1333).
// FLAC decoding is a serial walk over a bit stream (Rice-coded residuals, a linear-prediction recurrence per channel):
// there is nothing for 148 SMs to do, so this runs on a host core at a few hundred MB/s of PCM — a 60 s stereo clip
// decodes in tens of ms, once per export — and the samples then take the same road to HBM as every other clip.
// Follows the format as published (RFC 9639): STREAMINFO, frame header (UTF-8 coded number, CRC-8), CONSTANT /
// VERBATIM / FIXED (order 0-4) / LPC (order 1-32) subframes, wasted bits, partitioned Rice residuals with escape codes,
// left-side / side-right / mid-side stereo, CRC-16 per frame. Every CRC is checked; the MD5 of STREAMINFO is returned
// for the caller to check against the decoded samples (audio/reader.py does). No GPU, no CUDA call.
#include "sfb_internal.h"

#include <stdint.h>
#include <vector>

namespace {

struct Bits {
    const uint8_t* p; size_t n, pos = 0; uint64_t buf = 0; int have = 0; bool bad = false;
    Bits(const uint8_t* data, size_t bytes, size_t at) : p(data), n(bytes), pos(at) {}
    bool fill(int want) {
        while (have < want) {
            if (pos >= n) { bad = true; return false; }
            buf = (buf << 8) | p[pos++]; have += 8;
        }
        return true;
    }
    uint32_t read(int bits) {                                   // 0 <= bits <= 32
        if (bits == 0) return 0;
        if (!fill(bits)) return 0;
        const uint64_t v = (buf >> (have - bits)) & ((bits == 64) ? ~0ull : ((1ull << bits) - 1));
        have -= bits;
        return uint32_t(v);
    }
    int64_t read_signed(int bits) {                             // two's complement, 1 <= bits <= 33
        if (bits <= 32) {
            const uint32_t v = read(bits);
            const uint32_t sign = 1u << (bits - 1);
            return int64_t(v ^ sign) - int64_t(sign);
        }
        const uint64_t hi = read(bits - 32), lo = read(32);
        const uint64_t v = (hi << 32) | lo, sign = 1ull << (bits - 1);
        return int64_t(v ^ sign) - int64_t(sign);
    }
    uint32_t unary() {                                          // zeros before the next 1 bit
        uint32_t q = 0;
        for (;;) {
            if (have == 0 && !fill(8)) return q;
            const uint64_t window = buf & ((1ull << have) - 1);
            if (window == 0) { q += uint32_t(have); have = 0; continue; }
            const int top = 63 - __builtin_clzll(window);       // position of the first 1 among the `have` low bits
            q += uint32_t(have - 1 - top);
            have = top;
            return q;
        }
    }
    void align() { have -= have % 8; }
    size_t byte_pos() const { return pos - size_t(have/8); }
};

uint8_t crc8(const uint8_t* p, size_t n) {
    uint8_t c = 0;
    for (size_t i = 0; i < n; i++) { c ^= p[i]; for (int k = 0; k < 8; k++) c = (c & 0x80) ? uint8_t((c << 1) ^ 0x07) : uint8_t(c << 1); }
    return c;
}
uint16_t crc16(const uint8_t* p, size_t n) {
    static uint16_t table[256]; static bool ready = false;
    if (!ready) {
        for (int i = 0; i < 256; i++) { uint16_t c = uint16_t(i << 8); for (int k = 0; k < 8; k++) c = (c & 0x8000) ? uint16_t((c << 1) ^ 0x8005) : uint16_t(c << 1); table[i] = c; }
        ready = true;
    }
    uint16_t c = 0;
    for (size_t i = 0; i < n; i++) c = uint16_t((c << 8) ^ table[(c >> 8) ^ p[i]]);
    return c;
}

struct Stream { sfb_flac_info info; size_t audio; };

int parse_header(const uint8_t* d, size_t n, Stream& s) {
    size_t at = 0;
    if (n >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3')                         // an ID3v2 tag in front
        at = 10 + ((size_t(d[6] & 0x7f) << 21) | (size_t(d[7] & 0x7f) << 14) | (size_t(d[8] & 0x7f) << 7) | size_t(d[9] & 0x7f));
    SFB_REQUIRE(at + 4 <= n && d[at] == 'f' && d[at + 1] == 'L' && d[at + 2] == 'a' && d[at + 3] == 'C', "not a FLAC stream (no fLaC marker)");
    at += 4;
    bool seen = false;
    for (;;) {
        SFB_REQUIRE(at + 4 <= n, "FLAC: truncated metadata");
        const bool last = (d[at] & 0x80) != 0; const int type = d[at] & 0x7f;
        const size_t len = (size_t(d[at + 1]) << 16) | (size_t(d[at + 2]) << 8) | size_t(d[at + 3]);
        at += 4;
        SFB_REQUIRE(at + len <= n, "FLAC: truncated metadata block");
        if (type == 0) {
            SFB_REQUIRE(len >= 34, "FLAC: short STREAMINFO");
            const uint8_t* b = d + at;
            s.info.min_block = (b[0] << 8) | b[1]; s.info.max_block = (b[2] << 8) | b[3];
            s.info.samplerate = (int32_t(b[10]) << 12) | (int32_t(b[11]) << 4) | (b[12] >> 4);
            s.info.channels = ((b[12] >> 1) & 7) + 1;
            s.info.bits_per_sample = (((b[12] & 1) << 4) | (b[13] >> 4)) + 1;
            s.info.total_samples = (int64_t(b[13] & 15) << 32) | (int64_t(b[14]) << 24) | (int64_t(b[15]) << 16) | (int64_t(b[16]) << 8) | int64_t(b[17]);
            s.info.has_md5 = 0;
            for (int i = 0; i < 16; i++) { s.info.md5[i] = b[18 + i]; if (b[18 + i]) s.info.has_md5 = 1; }
            seen = true;
        }
        at += len;
        if (last) break;
    }
    SFB_REQUIRE(seen, "FLAC: no STREAMINFO block");
    SFB_REQUIRE(s.info.samplerate > 0 && s.info.bits_per_sample >= 4 && s.info.bits_per_sample <= 32, "FLAC: bad STREAMINFO");
    s.audio = at;
    return SFB_OK;
}

int residual(Bits& in, int64_t* out, int blocksize, int order) {
    const int method = int(in.read(2));
    SFB_REQUIRE(method < 2, "FLAC: reserved residual coding method");
    const int pbits = method ? 5 : 4, escape = method ? 31 : 15;
    const int porder = int(in.read(4)), parts = 1 << porder;
    SFB_REQUIRE((blocksize % parts) == 0 && (blocksize >> porder) >= order, "FLAC: partition order %d does not fit block %d / order %d", porder, blocksize, order);
    int at = order;
    for (int part = 0; part < parts; part++) {
        const int count = (blocksize >> porder) - (part ? 0 : order);
        const int k = int(in.read(pbits));
        if (k == escape) {
            const int raw = int(in.read(5));
            for (int i = 0; i < count; i++) out[at++] = raw ? in.read_signed(raw) : 0;
        } else {
            for (int i = 0; i < count; i++) {
                const uint64_t u = (uint64_t(in.unary()) << k) | in.read(k);
                out[at++] = int64_t(u >> 1) ^ -int64_t(u & 1);
            }
        }
        SFB_REQUIRE(!in.bad, "FLAC: stream ends inside a residual");
    }
    return SFB_OK;
}

int subframe(Bits& in, int64_t* out, int blocksize, int bps) {
    SFB_REQUIRE(in.read(1) == 0, "FLAC: subframe padding bit set");
    const int type = int(in.read(6));
    int wasted = 0;
    if (in.read(1)) wasted = int(in.unary()) + 1;
    bps -= wasted;
    SFB_REQUIRE(bps >= 1, "FLAC: more wasted bits than sample bits");
    if (type == 0) {
        const int64_t v = in.read_signed(bps);
        for (int i = 0; i < blocksize; i++) out[i] = v;
    } else if (type == 1) {
        for (int i = 0; i < blocksize; i++) out[i] = in.read_signed(bps);
    } else if (type >= 8 && type <= 12) {
        const int order = type - 8;
        SFB_REQUIRE(order <= blocksize, "FLAC: fixed order exceeds the block");
        for (int i = 0; i < order; i++) out[i] = in.read_signed(bps);
        if (int e = residual(in, out, blocksize, order)) return e;
        for (int i = order; i < blocksize; i++) {
            int64_t p = 0;
            switch (order) {
                case 1: p = out[i - 1]; break;
                case 2: p = 2*out[i - 1] - out[i - 2]; break;
                case 3: p = 3*out[i - 1] - 3*out[i - 2] + out[i - 3]; break;
                case 4: p = 4*out[i - 1] - 6*out[i - 2] + 4*out[i - 3] - out[i - 4]; break;
            }
            out[i] += p;
        }
    } else if (type >= 32) {
        const int order = (type & 31) + 1;
        SFB_REQUIRE(order <= blocksize, "FLAC: LPC order exceeds the block");
        for (int i = 0; i < order; i++) out[i] = in.read_signed(bps);
        const int precision = int(in.read(4)) + 1;
        SFB_REQUIRE(precision != 16, "FLAC: reserved LPC precision");
        const int shift = int(in.read_signed(5));
        SFB_REQUIRE(shift >= 0, "FLAC: negative LPC shift");
        int64_t coef[32];
        for (int i = 0; i < order; i++) coef[i] = in.read_signed(precision);
        if (int e = residual(in, out, blocksize, order)) return e;
        for (int i = order; i < blocksize; i++) {
            int64_t acc = 0;
            for (int j = 0; j < order; j++) acc += coef[j]*out[i - 1 - j];
            out[i] += acc >> shift;
        }
    } else {
        SFB_FAIL(SFB_EINVAL, "FLAC: reserved subframe type %d", type);
    }
    if (wasted) for (int i = 0; i < blocksize; i++) out[i] = int64_t(uint64_t(out[i]) << wasted);
    SFB_REQUIRE(!in.bad, "FLAC: stream ends inside a subframe");
    return SFB_OK;
}

// One frame starting at byte `at` → its block size; samples appended to pcm (interleaved) when pcm != NULL
int frame(const uint8_t* d, size_t n, size_t at, const sfb_flac_info& info, std::vector<int64_t>& scratch,
          int32_t* pcm, int64_t room, int* blocksize, size_t* next) {
    Bits in(d, n, at);
    SFB_REQUIRE(in.read(14) == 0x3FFE, "FLAC: lost frame sync at byte %zu", at);
    SFB_REQUIRE(in.read(1) == 0, "FLAC: reserved header bit set");
    in.read(1);                                                  // blocking strategy: the coded number is not needed here
    const int bs_code = int(in.read(4)), sr_code = int(in.read(4)), assignment = int(in.read(4)), ss_code = int(in.read(3));
    SFB_REQUIRE(in.read(1) == 0, "FLAC: reserved header bit set");
    const uint32_t first = in.read(8);                           // UTF-8 coded frame / sample number
    int extra = 0;
    if (first >= 0x80) { uint32_t m = first; while (m & 0x80) { extra++; m <<= 1; } extra--; SFB_REQUIRE(extra >= 1 && extra <= 6, "FLAC: bad coded number"); }
    for (int i = 0; i < extra; i++) SFB_REQUIRE((in.read(8) & 0xC0) == 0x80, "FLAC: bad coded number");
    int bs = 0;
    if (bs_code == 1) bs = 192;
    else if (bs_code >= 2 && bs_code <= 5) bs = 576 << (bs_code - 2);
    else if (bs_code == 6) bs = int(in.read(8)) + 1;
    else if (bs_code == 7) bs = int(in.read(16)) + 1;
    else if (bs_code >= 8) bs = 256 << (bs_code - 8);
    SFB_REQUIRE(bs > 0, "FLAC: reserved block size code");
    if (sr_code == 12) in.read(8); else if (sr_code == 13 || sr_code == 14) in.read(16);
    SFB_REQUIRE(sr_code != 15, "FLAC: invalid sample rate code");
    SFB_REQUIRE(!in.bad, "FLAC: stream ends inside a frame header");
    const size_t header_end = in.byte_pos();
    SFB_REQUIRE(in.read(8) == crc8(d + at, header_end - at), "FLAC: frame header CRC-8 mismatch at byte %zu", at);
    static const int sizes[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    const int bps = ss_code ? sizes[ss_code] : info.bits_per_sample;
    SFB_REQUIRE(bps > 0, "FLAC: reserved sample size code");
    const int channels = (assignment < 8) ? assignment + 1 : 2;
    SFB_REQUIRE(assignment <= 10, "FLAC: reserved channel assignment");
    SFB_REQUIRE(channels == info.channels, "FLAC: frame with %d channels in a %d-channel stream", channels, info.channels);
    scratch.resize(size_t(channels)*size_t(bs));
    for (int c = 0; c < channels; c++) {
        const bool side = (assignment == 8 && c == 1) || (assignment == 9 && c == 0) || (assignment == 10 && c == 1);
        if (int e = subframe(in, scratch.data() + size_t(c)*size_t(bs), bs, bps + (side ? 1 : 0))) return e;
    }
    in.align();
    const size_t body_end = in.byte_pos();
    const uint32_t stored = in.read(16);
    SFB_REQUIRE(!in.bad, "FLAC: stream ends inside a frame");
    SFB_REQUIRE(stored == crc16(d + at, body_end - at), "FLAC: frame CRC-16 mismatch at byte %zu", at);
    int64_t* a = scratch.data(); int64_t* b = scratch.data() + bs;
    if (assignment == 8) { for (int i = 0; i < bs; i++) b[i] = a[i] - b[i]; }                       // left, side → right
    else if (assignment == 9) { for (int i = 0; i < bs; i++) a[i] = a[i] + b[i]; }                  // side, right → left
    else if (assignment == 10) {
        for (int i = 0; i < bs; i++) { const int64_t side = b[i], mid = (a[i] << 1) | (side & 1); a[i] = (mid + side) >> 1; b[i] = (mid - side) >> 1; }
    }
    if (pcm) {
        SFB_REQUIRE(room >= bs, "FLAC: the stream holds more samples than the destination (%lld left, block of %d)", (long long)room, bs);
        for (int i = 0; i < bs; i++) for (int c = 0; c < channels; c++) pcm[size_t(i)*size_t(channels) + c] = int32_t(scratch[size_t(c)*size_t(bs) + i]);
    }
    *blocksize = bs; *next = in.byte_pos();
    return SFB_OK;
}

}  // namespace

extern "C" int sfb_flac_info_get(const void* data, size_t bytes, sfb_flac_info* info) {
    SFB_REQUIRE(data && info, "sfb_flac_info_get: null argument");
    Stream s{};
    if (int e = parse_header(static_cast<const uint8_t*>(data), bytes, s)) return e;
    *info = s.info;
    return SFB_OK;
}

extern "C" int sfb_flac_decode(const void* data, size_t bytes, int32_t* pcm, int64_t capacity_frames, int64_t* decoded_frames) {
    SFB_REQUIRE(data && decoded_frames, "sfb_flac_decode: null argument");
    const uint8_t* d = static_cast<const uint8_t*>(data);
    Stream s{};
    if (int e = parse_header(d, bytes, s)) return e;
    std::vector<int64_t> scratch;
    size_t at = s.audio;
    int64_t total = 0;
    while (at + 2 <= bytes) {
        if (!(d[at] == 0xFF && (d[at + 1] & 0xFE) == 0xF8)) break;      // trailing bytes that are not a frame (tags) end the audio
        int bs = 0; size_t next = at;
        if (int e = frame(d, bytes, at, s.info, scratch, pcm ? pcm + size_t(total)*size_t(s.info.channels) : nullptr,
                          capacity_frames - total, &bs, &next)) return e;
        total += bs; at = next;
    }
    *decoded_frames = total;
    return SFB_OK;
}
