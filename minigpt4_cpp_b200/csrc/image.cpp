// image.cpp — see image.h.  Everything here is host work the reference also does on the host (OpenCV + pillow-resize, minigpt4.cpp:2576-2651);
// it feeds minigpt4_encode_image and is not part of the measured hot path.
#include "image.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace mg4 {

// ------------------------------------------------------------------------------------------------
// DEFLATE (RFC 1951) inside a zlib stream (RFC 1950)
// ------------------------------------------------------------------------------------------------
namespace {

struct Bits {
    const uint8_t *p; size_t n, pos = 0;
    uint64_t buf = 0; int cnt = 0; size_t overrun = 0;   // bytes read past the end (as zeros): an error if any of them is ever consumed
    void fill() { while (cnt <= 56) { uint64_t b = 0; if (pos < n) b = p[pos]; else ++overrun; ++pos; buf |= b << cnt; cnt += 8; } }
    uint32_t peek(int k) { if (cnt < k) fill(); return (uint32_t)(buf & ((1ull << k) - 1)); }
    void drop(int k) { buf >>= k; cnt -= k; }
    uint32_t get(int k) { if (k == 0) return 0; const uint32_t v = peek(k); drop(k); return v; }
    bool past_end() const { return pos - (size_t)(cnt / 8) > n; }   // more whole bytes consumed than the stream holds
    void align_byte() { drop(cnt & 7); }
};

constexpr int kFast = 10;
struct Huff {
    uint16_t count[16]; uint16_t sym[320];
    uint16_t fast[1 << kFast];   // (length << 9) | symbol for codes of <= kFast bits, 0 = take the slow path
    bool build(const uint8_t *len, int n) {
        memset(count, 0, sizeof count); memset(fast, 0, sizeof fast);
        for (int i = 0; i < n; ++i) ++count[len[i]];
        count[0] = 0;
        int left = 1;
        for (int l = 1; l < 16; ++l) { left <<= 1; left -= count[l]; if (left < 0) return false; }   // over-subscribed
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; ++i) if (len[i]) sym[offs[len[i]]++] = (uint16_t)i;
        // canonical codes, most significant bit first; the stream carries them least significant bit first
        int code = 0, idx = 0;
        for (int l = 1; l <= kFast; ++l) {
            for (int k = 0; k < count[l]; ++k, ++code, ++idx) {
                int rev = 0;
                for (int b = 0; b < l; ++b) if (code & (1 << b)) rev |= 1 << (l - 1 - b);
                for (int hi = rev; hi < (1 << kFast); hi += 1 << l) fast[hi] = (uint16_t)((l << 9) | sym[idx]);
            }
            code <<= 1;
        }
        return true;
    }
    int decode(Bits &br) const {
        const uint16_t e = fast[br.peek(kFast)];
        if (e) { br.drop(e >> 9); return e & 511; }
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; ++l) {
            code |= (int)br.get(1);
            const int c = count[l];
            if (code - c < first) return sym[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

uint32_t adler32(const uint8_t *d, size_t n) {
    uint32_t a = 1, b = 0;
    while (n) {
        const size_t k = n < 5552 ? n : 5552;
        for (size_t i = 0; i < k; ++i) { a += d[i]; b += a; }
        a %= 65521; b %= 65521; d += k; n -= k;
    }
    return (b << 16) | a;
}

}  // namespace

bool inflate_zlib(const uint8_t *src, size_t n, std::vector<uint8_t> &dst, size_t expected, std::string &err) {
    if (n < 6) { err = "zlib stream too short"; return false; }
    if ((src[0] & 15) != 8 || ((src[0] << 8) | src[1]) % 31 || (src[1] & 32)) { err = "not a deflate zlib stream"; return false; }
    Bits br{src + 2, n - 2};
    dst.clear(); dst.reserve(expected);
    Huff lit, dist;
    for (bool last = false; !last;) {
        last = br.get(1) != 0;
        const uint32_t type = br.get(2);
        if (type == 0) {
            br.align_byte();
            const uint32_t len = br.get(16), nlen = br.get(16);
            if ((len ^ nlen) != 0xFFFFu) { err = "stored block length check failed"; return false; }
            for (uint32_t i = 0; i < len; ++i) dst.push_back((uint8_t)br.get(8));
            if (br.past_end()) { err = "deflate stream truncated"; return false; }
            continue;
        }
        if (type == 3) { err = "reserved deflate block type"; return false; }
        uint8_t lens[320];
        if (type == 1) {
            for (int i = 0; i < 144; ++i) lens[i] = 8;
            for (int i = 144; i < 256; ++i) lens[i] = 9;
            for (int i = 256; i < 280; ++i) lens[i] = 7;
            for (int i = 280; i < 288; ++i) lens[i] = 8;
            lit.build(lens, 288);
            for (int i = 0; i < 30; ++i) lens[i] = 5;
            dist.build(lens, 30);
        } else {
            const int hlit = (int)br.get(5) + 257, hdist = (int)br.get(5) + 1, hclen = (int)br.get(4) + 4;
            if (hlit > 286 || hdist > 30) { err = "bad dynamic block header"; return false; }
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; ++i) cl[kClOrder[i]] = (uint8_t)br.get(3);
            Huff clh;
            if (!clh.build(cl, 19)) { err = "bad code-length code"; return false; }
            int i = 0;
            while (i < hlit + hdist) {
                const int s = clh.decode(br);
                if (s < 0) { err = "bad code-length symbol"; return false; }
                if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                int rep, val = 0;
                if (s == 16) { if (i == 0) { err = "repeat without a previous length"; return false; } val = lens[i - 1]; rep = 3 + (int)br.get(2); }
                else if (s == 17) rep = 3 + (int)br.get(3);
                else rep = 11 + (int)br.get(7);
                if (i + rep > hlit + hdist) { err = "code lengths overflow"; return false; }
                while (rep--) lens[i++] = (uint8_t)val;
            }
            if (lens[256] == 0) { err = "no end-of-block code"; return false; }
            if (!lit.build(lens, hlit) || !dist.build(lens + hlit, hdist)) { err = "over-subscribed Huffman code"; return false; }
        }
        for (;;) {
            const int s = lit.decode(br);
            if (s < 0) { err = "bad literal/length symbol"; return false; }
            if (s < 256) {
                dst.push_back((uint8_t)s);
                if ((dst.size() & 4095) == 0 && (br.past_end() || dst.size() > expected + (1u << 16))) { err = "deflate stream truncated or longer than the image needs"; return false; }
                continue;
            }
            if (s == 256) break;
            if (s > 285) { err = "bad length symbol"; return false; }
            const int len = kLenBase[s - 257] + (int)br.get(kLenExtra[s - 257]);
            const int ds = dist.decode(br);
            if (ds < 0 || ds > 29) { err = "bad distance symbol"; return false; }
            const size_t d = (size_t)kDistBase[ds] + br.get(kDistExtra[ds]);
            if (d > dst.size()) { err = "distance reaches before the start of the output"; return false; }
            const size_t at = dst.size();
            dst.resize(at + (size_t)len);
            for (int i = 0; i < len; ++i) dst[at + (size_t)i] = dst[at + (size_t)i - d];   // (may overlap: byte by byte)
            if (br.past_end()) { err = "deflate stream truncated"; return false; }
            if (dst.size() > expected + (1u << 16)) { err = "deflate output larger than the image needs"; return false; }
        }
        if (br.past_end()) { err = "deflate stream truncated"; return false; }
    }
    br.align_byte();
    uint32_t want = 0;
    for (int i = 0; i < 4; ++i) want = (want << 8) | br.get(8);
    if (br.past_end()) { err = "zlib stream has no checksum"; return false; }
    if (want != adler32(dst.data(), dst.size())) { err = "zlib checksum mismatch"; return false; }
    return true;
}

// ------------------------------------------------------------------------------------------------
// PNG (ISO/IEC 15948)
// ------------------------------------------------------------------------------------------------
namespace {

struct CrcTable {
    uint32_t t[256];
    CrcTable() { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; t[i] = c; } }
};
uint32_t crc32_png(const uint8_t *d, size_t n) {
    static const CrcTable table;   // (initialised once, thread-safe)
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table.t[(c ^ d[i]) & 255] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

bool decode_png(const uint8_t *data, size_t n, RgbImage &out, std::string &err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8 || memcmp(data, sig, 8)) { err = "not a PNG file"; return false; }
    size_t pos = 8;
    uint32_t w = 0, h = 0; int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat; uint8_t plte[768]; int n_plte = 0; bool have_ihdr = false, have_iend = false;
    while (pos + 12 <= n && !have_iend) {
        const uint32_t len = be32(data + pos);
        if (len > n - pos - 12) { err = "PNG chunk runs past the end of the file"; return false; }
        const uint8_t *type = data + pos + 4, *body = data + pos + 8;
        if (crc32_png(type, (size_t)len + 4) != be32(body + len)) { err = "PNG chunk checksum mismatch"; return false; }
        if (!have_ihdr && memcmp(type, "IHDR", 4)) { err = "PNG does not start with IHDR"; return false; }
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13 || have_ihdr) { err = "bad IHDR"; return false; }
            w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            if (body[10] != 0 || body[11] != 0 || interlace > 1) { err = "unsupported PNG compression / filter / interlace method"; return false; }
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            if (len % 3 || len > 768) { err = "bad PLTE"; return false; }
            memcpy(plte, body, len); n_plte = (int)len / 3;
        } else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) have_iend = true;
        else if (!(type[0] & 32)) { err = "unknown critical PNG chunk"; return false; }
        pos += (size_t)len + 12;
    }
    if (!have_ihdr || !have_iend) { err = "truncated PNG"; return false; }
    int samples;
    switch (ctype) {
        case 0: samples = 1; if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) samples = 0; break;
        case 2: samples = 3; if (depth != 8 && depth != 16) samples = 0; break;
        case 3: samples = 1; if (depth != 1 && depth != 2 && depth != 4 && depth != 8) samples = 0; if (!n_plte) { err = "palette image without PLTE"; return false; } break;
        case 4: samples = 2; if (depth != 8 && depth != 16) samples = 0; break;
        case 6: samples = 4; if (depth != 8 && depth != 16) samples = 0; break;
        default: samples = 0;
    }
    if (!samples) { err = "invalid PNG colour type / bit depth"; return false; }
    if (w == 0 || h == 0 || w > 32768 || h > 32768 || (uint64_t)w * h > (1ull << 28)) { err = "PNG dimensions out of range"; return false; }
    const int bits = samples * depth, bpp = std::max(1, bits / 8);
    // passes: {x0, y0, dx, dy}; a non-interlaced image is one pass
    static const int adam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const int whole[1][4] = {{0, 0, 1, 1}};
    const int (*pass)[4] = interlace ? adam7 : whole; const int n_pass = interlace ? 7 : 1;
    size_t expected = 0;
    for (int p = 0; p < n_pass; ++p) {
        const size_t pw = (w > (uint32_t)pass[p][0]) ? (w - pass[p][0] + pass[p][2] - 1) / pass[p][2] : 0, ph = (h > (uint32_t)pass[p][1]) ? (h - pass[p][1] + pass[p][3] - 1) / pass[p][3] : 0;
        if (pw && ph) expected += ph * (1 + (pw * (size_t)bits + 7) / 8);
    }
    std::vector<uint8_t> raw;
    if (!inflate_zlib(idat.data(), idat.size(), raw, expected, err)) return false;
    if (raw.size() < expected) { err = "PNG image data too short"; return false; }
    out.w = (int)w; out.h = (int)h; out.px.assign((size_t)w * h * 3, 0);
    size_t at = 0;
    std::vector<uint8_t> prev, cur;
    for (int p = 0; p < n_pass; ++p) {
        const size_t pw = (w > (uint32_t)pass[p][0]) ? (w - pass[p][0] + pass[p][2] - 1) / pass[p][2] : 0, ph = (h > (uint32_t)pass[p][1]) ? (h - pass[p][1] + pass[p][3] - 1) / pass[p][3] : 0;
        if (!pw || !ph) continue;
        const size_t rb = (pw * (size_t)bits + 7) / 8;
        prev.assign(rb, 0); cur.resize(rb);
        for (size_t y = 0; y < ph; ++y) {
            const int ft = raw[at++];
            const uint8_t *s = raw.data() + at; at += rb;
            if (ft > 4) { err = "bad PNG filter type"; return false; }
            for (size_t i = 0; i < rb; ++i) {
                const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
                int v = s[i];
                switch (ft) { case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break; case 4: v += paeth(a, b, c); break; default: break; }
                cur[i] = (uint8_t)v;
            }
            uint8_t *row = out.px.data() + ((size_t)(pass[p][1] + y * pass[p][3]) * w) * 3;
            for (size_t x = 0; x < pw; ++x) {
                uint8_t sv[4];
                for (int k = 0; k < samples; ++k) {
                    const size_t si = x * samples + k;
                    if (depth == 8) sv[k] = cur[si];
                    else if (depth == 16) sv[k] = cur[2 * si];   // high byte
                    else { const size_t bit = si * depth; sv[k] = (uint8_t)((cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1)); }
                }
                uint8_t *px = row + (size_t)(pass[p][0] + x * pass[p][2]) * 3;
                if (ctype == 3) {
                    const int idx = sv[0];
                    if (idx < n_plte) { px[0] = plte[3 * idx]; px[1] = plte[3 * idx + 1]; px[2] = plte[3 * idx + 2]; }
                } else if (ctype == 0 || ctype == 4) {
                    const uint8_t g = depth < 8 ? (uint8_t)(sv[0] * 255 / ((1 << depth) - 1)) : sv[0];
                    px[0] = px[1] = px[2] = g;
                } else { px[0] = sv[0]; px[1] = sv[1]; px[2] = sv[2]; }
            }
            prev.swap(cur);
        }
    }
    return true;
}

// binary PPM (P6) / PGM (P5), 8 bit
static bool decode_pnm(const uint8_t *d, size_t n, RgbImage &out, std::string &err) {
    size_t pos = 2; long v[3]; int got = 0;
    while (got < 3 && pos < n) {
        if (d[pos] == '#') { while (pos < n && d[pos] != '\n') ++pos; continue; }
        if (d[pos] == ' ' || d[pos] == '\t' || d[pos] == '\r' || d[pos] == '\n') { ++pos; continue; }
        if (d[pos] < '0' || d[pos] > '9') { err = "bad PNM header"; return false; }
        long x = 0; while (pos < n && d[pos] >= '0' && d[pos] <= '9' && x < (1 << 28)) x = x * 10 + (d[pos++] - '0');
        v[got++] = x;
    }
    if (got < 3 || pos >= n) { err = "truncated PNM header"; return false; }
    ++pos;   // the single whitespace byte after maxval
    const int ch = d[1] == '6' ? 3 : 1;
    if (v[0] <= 0 || v[1] <= 0 || v[0] > 32768 || v[1] > 32768 || v[2] <= 0 || v[2] > 255) { err = "unsupported PNM dimensions / maxval"; return false; }
    const size_t need = (size_t)v[0] * v[1] * ch;
    if (n - pos < need) { err = "truncated PNM data"; return false; }
    out.w = (int)v[0]; out.h = (int)v[1]; out.px.resize((size_t)out.w * out.h * 3);
    for (size_t i = 0; i < (size_t)out.w * out.h; ++i)
        for (int k = 0; k < 3; ++k) out.px[3 * i + k] = d[pos + i * ch + (ch == 3 ? k : 0)];
    return true;
}

bool decode_image_file(const char *path, RgbImage &out, std::string &err) {
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) { err = "cannot open file"; return false; }
    std::vector<uint8_t> buf;
    uint8_t tmp[65536]; size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) { buf.insert(buf.end(), tmp, tmp + k); if (buf.size() > (1ull << 30)) break; }
    fclose(f);
    if (buf.size() >= 8 && buf[0] == 0x89 && buf[1] == 'P') return decode_png(buf.data(), buf.size(), out, err);
    if (buf.size() >= 8 && buf[0] == 'P' && (buf[1] == '6' || buf[1] == '5')) return decode_pnm(buf.data(), buf.size(), out, err);
    if (buf.size() >= 3 && buf[0] == 0xFF && buf[1] == 0xD8) return decode_jpeg(buf.data(), buf.size(), out, err);
    err = "unrecognised image format (PNG, JPEG, PPM, PGM are decoded)";
    return false;
}

// ------------------------------------------------------------------------------------------------
// Pillow's bicubic resize, 8 bits per channel
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

inline double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

struct Coeffs { int ksize = 0; std::vector<int> bounds; std::vector<int32_t> kk; };

Coeffs precompute(int in_size, int out_size) {
    Coeffs c;
    const float in0 = 0.f, in1 = (float)in_size;
    double filterscale, scale;
    filterscale = scale = (double)(in1 - in0) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;   // bicubic support 2
    c.ksize = (int)ceil(support) * 2 + 1;
    c.bounds.resize((size_t)out_size * 2); c.kk.assign((size_t)out_size * c.ksize, 0);
    std::vector<double> k((size_t)c.ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) { const double w = bicubic_filter((x + xmin - center + 0.5) * ss); k[(size_t)x] = w; ww += w; }
        for (int x = 0; x < xmax; ++x) {
            double v = k[(size_t)x];
            if (ww != 0.0) v /= ww;
            c.kk[(size_t)xx * c.ksize + x] = v < 0 ? (int32_t)(-0.5 + v * (1 << kPrecisionBits)) : (int32_t)(0.5 + v * (1 << kPrecisionBits));
        }
        c.bounds[(size_t)xx * 2] = xmin; c.bounds[(size_t)xx * 2 + 1] = xmax;
    }
    return c;
}
inline uint8_t clip8(int32_t v) { v >>= kPrecisionBits; return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

}  // namespace

void resize_bicubic_u8(const uint8_t *src, int w, int h, uint8_t *dst, int ow, int oh) {
    if (w == ow && h == oh) { memcpy(dst, src, (size_t)w * h * 3); return; }
    const uint8_t *cur = src; int cw = w;
    std::vector<uint8_t> tmp;
    Coeffs cv;
    int y_first = 0, y_last = h;
    if (h != oh) {   // only the source rows the vertical pass reads go through the horizontal pass
        cv = precompute(h, oh);
        y_first = cv.bounds[0]; y_last = cv.bounds[(size_t)oh * 2 - 2] + cv.bounds[(size_t)oh * 2 - 1];
    }
    if (w != ow) {
        const Coeffs ch = precompute(w, ow);
        tmp.assign((size_t)h * ow * 3, 0);
        uint8_t *t = tmp.data();
        parallel_rows(y_last - y_first, (size_t)ow * ch.ksize * 3, [&](int a, int b) {
            for (int yy = y_first + a; yy < y_first + b; ++yy) {
                const uint8_t *row = src + (size_t)yy * w * 3;
                for (int xx = 0; xx < ow; ++xx) {
                    const int xmin = ch.bounds[(size_t)xx * 2], xmax = ch.bounds[(size_t)xx * 2 + 1];
                    const int32_t *k = &ch.kk[(size_t)xx * ch.ksize];
                    int32_t s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
                    for (int x = 0; x < xmax; ++x) { const uint8_t *p = row + (size_t)(x + xmin) * 3; s0 += p[0] * k[x]; s1 += p[1] * k[x]; s2 += p[2] * k[x]; }
                    uint8_t *o = t + ((size_t)yy * ow + xx) * 3;
                    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
                }
            }
        });
        cur = t; cw = ow;
    }
    if (h != oh) {
        parallel_rows(oh, (size_t)cw * cv.ksize * 3, [&](int a, int b) {
            for (int yy = a; yy < b; ++yy) {
                const int ymin = cv.bounds[(size_t)yy * 2], ymax = cv.bounds[(size_t)yy * 2 + 1];
                const int32_t *k = &cv.kk[(size_t)yy * cv.ksize];
                for (int xx = 0; xx < cw * 3; ++xx) {
                    int32_t s = 1 << (kPrecisionBits - 1);
                    for (int y = 0; y < ymax; ++y) s += cur[(size_t)(y + ymin) * cw * 3 + xx] * k[y];
                    dst[(size_t)yy * cw * 3 + xx] = clip8(s);
                }
            }
        });
    } else memcpy(dst, cur, (size_t)cw * h * 3);
}

void normalize_to_chw(const uint8_t *rgb, int w, int h, float *out) {
    static const double mean[3] = {0.48145466, 0.4578275, 0.40821073}, sd[3] = {0.26862954, 0.26130258, 0.27577711};
    const float inv255 = 1.0f / 255.0f;
    const size_t plane = (size_t)w * h;
    for (size_t i = 0; i < plane; ++i)
        for (int c = 0; c < 3; ++c) {
            const float f = (float)rgb[3 * i + c] * inv255;
            const float d = (float)((double)f - mean[c]);
            out[(size_t)c * plane + i] = (float)((double)d / sd[c]);
        }
}

}  // namespace mg4
