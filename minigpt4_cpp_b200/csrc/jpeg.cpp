// jpeg.cpp — JPEG (ITU-T T.81) decoder behind minigpt4_image_load_from_file: baseline, extended-sequential and progressive Huffman streams,
// 8-bit, greyscale or three components (YCbCr, or RGB when the file says so), any scan layout, restart intervals, EXIF orientation.
// The reference reads JPEG through cv::imread = libjpeg-turbo with its defaults; the stages whose arithmetic decides the pixels restate that
// library's published algorithms so that the bytes agree: the 13-bit "islow" inverse DCT (jidctint.c), the triangle-filter ("fancy") chroma
// upsampling of jdsample.c with its edge replication, and the 16-bit fixed-point YCbCr -> RGB tables of jdcolor.c.  Pillow decodes with the
// same library and defaults and is the checker in tests/test_image_cpu.py.  Not decoded: 12-bit, lossless, arithmetic coding, CMYK.
#include "image.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace mg4 {
namespace {

const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    uint8_t bits[17] = {0}, vals[256] = {0};
    int mincode[18], maxcode[18], valptr[18];
    uint16_t look[512];   // 9-bit prefix -> (length << 8) | symbol, 0 = longer code
    bool build() {
        int code = 0, k = 0;
        memset(look, 0, sizeof look);
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k; mincode[l] = code;
            if (code + bits[l] > (1 << l) || k + bits[l] > 256) return false;   // over-subscribed code
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                if (l <= 9) for (int f = 0; f < (1 << (9 - l)); ++f) look[(code << (9 - l)) | f] = (uint16_t)((l << 8) | vals[k]);
            }
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        return true;
    }
};

// entropy-coded segment reader: most significant bit first, FF 00 -> FF, any other FF xx is a marker that ends the segment (zeros are fed after it)
struct BitReader {
    const uint8_t *p; size_t n, pos;
    uint32_t buf = 0; int cnt = 0; int marker = 0; size_t fed_zero_bytes = 0;
    void fill() {
        while (cnt <= 24) {
            uint32_t b = 0;
            if (!marker && pos < n) {
                b = p[pos++];
                if (b == 0xFF) {
                    while (pos < n && p[pos] == 0xFF) ++pos;   // fill bytes
                    const int m = pos < n ? p[pos] : 0xD9;
                    if (m == 0) ++pos;
                    else { marker = m; ++pos; b = 0; ++fed_zero_bytes; }
                }
            } else ++fed_zero_bytes;
            buf |= b << (24 - cnt); cnt += 8;
        }
    }
    int peek(int k) { if (cnt < k) fill(); return (int)(buf >> (32 - k)); }
    void drop(int k) { buf <<= k; cnt -= k; }
    int get(int k) { if (!k) return 0; const int v = peek(k); drop(k); return v; }
    void reset() { buf = 0; cnt = 0; }
};
inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
inline int huff_decode(BitReader &br, const HuffTable &h) {
    const uint16_t e = h.look[br.peek(9)];
    if (e) { br.drop(e >> 8); return e & 255; }
    int code = br.peek(16);
    for (int l = 10; l <= 16; ++l) {
        const int c = code >> (16 - l);
        if (h.maxcode[l] >= 0 && c <= h.maxcode[l] && c >= h.mincode[l]) { br.drop(l); return h.vals[h.valptr[l] + c - h.mincode[l]]; }
    }
    return -1;
}

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int bw = 0, bh = 0;          // blocks across / down as stored (padded to whole MCUs)
    int cw = 0, chh = 0;         // blocks that carry image data: ceil(component size / 8)
    int dw = 0, dh = 0;          // "downsampled" size in samples
    std::vector<int16_t> coef;   // [bh][bw][64], natural order
    std::vector<uint8_t> plane;  // [bh * 8][bw * 8]
    int pred = 0;
};

// jidctint.c: accurate integer inverse DCT (CONST_BITS 13, PASS1_BITS 2), dequantisation folded in
inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
void idct_islow(const int16_t *in, const uint16_t *q, uint8_t *out, int stride) {
    constexpr int CB = 13, P1 = 2;
    constexpr int F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299, F1_847 = 15137,
                  F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
    int ws[64];
    for (int c = 0; c < 8; ++c) {
        const int16_t *ip = in + c; const uint16_t *qp = q + c; int *wp = ws + c;
        long z2 = (long)ip[16] * qp[16], z3 = (long)ip[48] * qp[48];
        long z1 = (z2 + z3) * F0_541;
        long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        z2 = (long)ip[0] * qp[0]; z3 = (long)ip[32] * qp[32];
        long tmp0 = (z2 + z3) * (1L << CB), tmp1 = (z2 - z3) * (1L << CB);
        const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = (long)ip[56] * qp[56]; tmp1 = (long)ip[40] * qp[40]; tmp2 = (long)ip[24] * qp[24]; tmp3 = (long)ip[8] * qp[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; long z4 = tmp1 + tmp3;
        const long z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const long r = 1L << (CB - P1 - 1);
        wp[0] = (int)((tmp10 + tmp3 + r) >> (CB - P1)); wp[56] = (int)((tmp10 - tmp3 + r) >> (CB - P1));
        wp[8] = (int)((tmp11 + tmp2 + r) >> (CB - P1)); wp[48] = (int)((tmp11 - tmp2 + r) >> (CB - P1));
        wp[16] = (int)((tmp12 + tmp1 + r) >> (CB - P1)); wp[40] = (int)((tmp12 - tmp1 + r) >> (CB - P1));
        wp[24] = (int)((tmp13 + tmp0 + r) >> (CB - P1)); wp[32] = (int)((tmp13 - tmp0 + r) >> (CB - P1));
    }
    for (int rr = 0; rr < 8; ++rr) {
        const int *wp = ws + rr * 8; uint8_t *op = out + (size_t)rr * stride;
        long z2 = wp[2], z3 = wp[6];
        long z1 = (z2 + z3) * F0_541;
        long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        long tmp0 = ((long)wp[0] + wp[4]) * (1L << CB), tmp1 = ((long)wp[0] - wp[4]) * (1L << CB);
        const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = wp[7]; tmp1 = wp[5]; tmp2 = wp[3]; tmp3 = wp[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; long z4 = tmp1 + tmp3;
        const long z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        constexpr int S = CB + P1 + 3; const long r = 1L << (S - 1);
        op[0] = clamp8((int)((tmp10 + tmp3 + r) >> S) + 128); op[7] = clamp8((int)((tmp10 - tmp3 + r) >> S) + 128);
        op[1] = clamp8((int)((tmp11 + tmp2 + r) >> S) + 128); op[6] = clamp8((int)((tmp11 - tmp2 + r) >> S) + 128);
        op[2] = clamp8((int)((tmp12 + tmp1 + r) >> S) + 128); op[5] = clamp8((int)((tmp12 - tmp1 + r) >> S) + 128);
        op[3] = clamp8((int)((tmp13 + tmp0 + r) >> S) + 128); op[4] = clamp8((int)((tmp13 - tmp0 + r) >> S) + 128);
    }
}

struct Decoder {
    const uint8_t *d; size_t n; std::string &err;
    int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1; bool progressive = false, have_sof = false;
    Component comp[3];
    uint16_t qt[4][64]; bool have_qt[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    int restart_interval = 0;
    bool jfif = false, adobe = false; int adobe_transform = 0; int orientation = 1;
    int eobrun = 0;

    bool fail(const char *m) { err = m; return false; }
    static int be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

    void parse_exif(const uint8_t *p, size_t len) {
        if (len < 14 || memcmp(p, "Exif\0\0", 6)) return;
        const uint8_t *t = p + 6; const size_t tl = len - 6;
        const bool le = t[0] == 'I' && t[1] == 'I';
        if (!le && !(t[0] == 'M' && t[1] == 'M')) return;
        auto r16 = [&](size_t o) -> uint32_t { return o + 2 <= tl ? (le ? t[o] | (t[o + 1] << 8) : (t[o] << 8) | t[o + 1]) : 0u; };
        auto r32 = [&](size_t o) -> uint32_t { return o + 4 <= tl ? (le ? r16(o) | (r16(o + 2) << 16) : (r16(o) << 16) | r16(o + 2)) : 0u; };
        if (r16(2) != 42) return;
        const size_t ifd = r32(4);
        const uint32_t cnt = r16(ifd);
        for (uint32_t i = 0; i < cnt && i < 512; ++i) {
            const size_t e = ifd + 2 + (size_t)i * 12;
            if (e + 12 > tl) return;
            if (r16(e) == 0x0112 && r16(e + 2) == 3) { const uint32_t v = r16(e + 8); if (v >= 1 && v <= 8) orientation = (int)v; return; }
        }
    }

    bool read_tables_and_header(size_t &pos, int &marker_out) {   // consumes markers up to and including an SOS or EOI header
        for (;;) {
            while (pos < n && d[pos] != 0xFF) ++pos;           // (garbage between segments is skipped like libjpeg's next_marker)
            while (pos < n && d[pos] == 0xFF) ++pos;
            if (pos >= n) return fail("JPEG ends without an end-of-image marker");
            const int m = d[pos++];
            if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (m == 0xD9) { marker_out = m; return true; }
            if (pos + 2 > n) return fail("truncated JPEG segment");
            const size_t len = (size_t)be16(d + pos);
            if (len < 2 || pos + len > n) return fail("truncated JPEG segment");
            const uint8_t *s = d + pos + 2; const size_t sl = len - 2;
            if (m == 0xDB) {   // DQT
                size_t o = 0;
                while (o < sl) {
                    const int pq = s[o] >> 4, tq = s[o] & 15; ++o;
                    if (tq > 3 || pq > 1 || o + (size_t)64 * (pq + 1) > sl) return fail("bad quantisation table");
                    for (int i = 0; i < 64; ++i) { qt[tq][kZigzag[i]] = (uint16_t)(pq ? be16(s + o + 2 * i) : s[o + i]); }
                    o += (size_t)64 * (pq + 1); have_qt[tq] = true;
                }
            } else if (m == 0xC4) {   // DHT
                size_t o = 0;
                while (o < sl) {
                    if (o + 17 > sl) return fail("bad Huffman table");
                    const int tc = s[o] >> 4, th = s[o] & 15;
                    if (tc > 1 || th > 3) return fail("bad Huffman table");
                    HuffTable &h = tc ? ac[th] : dc[th];
                    int total = 0; h.bits[0] = 0;
                    for (int l = 1; l <= 16; ++l) { h.bits[l] = s[o + l]; total += h.bits[l]; }
                    if (total > 256 || o + 17 + (size_t)total > sl) return fail("bad Huffman table");
                    memset(h.vals, 0, sizeof h.vals); memcpy(h.vals, s + o + 17, (size_t)total);
                    if (!h.build()) return fail("bad Huffman table");
                    h.present = true; o += 17 + (size_t)total;
                }
            } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
                if (have_sof) return fail("more than one frame header");
                if (sl < 6) return fail("bad frame header");
                if (s[0] != 8) return fail("only 8-bit JPEG is decoded");
                height = be16(s + 1); width = be16(s + 3); ncomp = s[5];
                if (ncomp == 4) return fail("CMYK / YCCK JPEG is not decoded");
                if ((ncomp != 1 && ncomp != 3) || sl < 6 + (size_t)3 * ncomp) return fail("bad frame header");
                if (width <= 0 || height <= 0 || width > 32768 || height > 32768 || (uint64_t)width * height > (1ull << 27)) return fail("JPEG dimensions out of range");
                for (int i = 0; i < ncomp; ++i) {
                    Component &c = comp[i];
                    c.id = s[6 + 3 * i]; c.h = s[7 + 3 * i] >> 4; c.v = s[7 + 3 * i] & 15; c.tq = s[8 + 3 * i];
                    if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("bad sampling factors");
                    hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v);
                }
                if (ncomp == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }   // (a single-component image has no interleave: the factors are moot)
                const int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
                for (int i = 0; i < ncomp; ++i) {
                    Component &c = comp[i];
                    if (hmax % c.h || vmax % c.v) return fail("fractional sampling ratios are not decoded");
                    c.dw = (width * c.h + hmax - 1) / hmax; c.dh = (height * c.v + vmax - 1) / vmax;
                    c.cw = (c.dw + 7) / 8; c.chh = (c.dh + 7) / 8;
                    c.bw = mcux * c.h; c.bh = mcuy * c.v;
                    c.coef.assign((size_t)c.bw * c.bh * 64, 0);
                }
                progressive = m == 0xC2; have_sof = true;
            } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8)) {
                return fail("lossless, hierarchical and arithmetic-coded JPEG are not decoded");
            } else if (m == 0xDD) {
                if (sl < 2) return fail("bad restart interval"); restart_interval = be16(s);
            } else if (m == 0xE0) { if (sl >= 5 && !memcmp(s, "JFIF", 5)) jfif = true; }
            else if (m == 0xE1) parse_exif(s, sl);
            else if (m == 0xEE) { if (sl >= 12 && !memcmp(s, "Adobe", 5)) { adobe = true; adobe_transform = s[11]; } }
            else if (m == 0xDA) { marker_out = m; return true; }   // (pos stays on the length field: decode_scan reads the header)
            pos += len;
        }
    }

    bool restart(BitReader &br, int &next_rst) {
        br.reset();
        if (!br.marker) {   // the marker has not been run into yet: find it
            size_t q = br.pos;
            while (q + 1 < n && !(d[q] == 0xFF && d[q + 1] != 0 && d[q + 1] != 0xFF)) ++q;
            if (q + 1 >= n) return fail("missing restart marker");
            br.marker = d[q + 1]; br.pos = q + 2;
        }
        if (br.marker != 0xD0 + next_rst) return fail("restart markers out of sequence");
        br.marker = 0; next_rst = (next_rst + 1) & 7;
        for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
        eobrun = 0;
        return true;
    }

    bool decode_block_baseline(BitReader &br, Component &c, const HuffTable &hd, const HuffTable &ha, int16_t *blk) {
        int s = huff_decode(br, hd);
        if (s < 0 || s > 15) return fail("bad DC code");
        const int diff = s ? extend(br.get(s), s) : 0;
        c.pred += diff; blk[0] = (int16_t)c.pred;
        for (int k = 1; k < 64;) {
            const int rs = huff_decode(br, ha);
            if (rs < 0) return fail("bad AC code");
            const int r = rs >> 4; s = rs & 15;
            if (!s) { if (r == 15) { k += 16; continue; } break; }
            k += r;
            if (k > 63) return fail("AC coefficient index out of range");
            blk[kZigzag[k]] = (int16_t)extend(br.get(s), s);
            ++k;
        }
        return true;
    }
    bool decode_block_progressive(BitReader &br, Component &c, const HuffTable *hd, const HuffTable *ha, int16_t *blk, int ss, int se, int ah, int al) {
        if (ss == 0) {
            if (ah == 0) {
                const int s = huff_decode(br, *hd);
                if (s < 0 || s > 15) return fail("bad DC code");
                c.pred += s ? extend(br.get(s), s) : 0;
                blk[0] = (int16_t)(c.pred * (1 << al));
            } else if (br.get(1)) blk[0] = (int16_t)(blk[0] | (1 << al));
            return true;
        }
        if (ah == 0) {
            if (eobrun > 0) { --eobrun; return true; }
            for (int k = ss; k <= se;) {
                const int rs = huff_decode(br, *ha);
                if (rs < 0) return fail("bad AC code");
                const int r = rs >> 4, s = rs & 15;
                if (!s) {
                    if (r == 15) { k += 16; continue; }
                    eobrun = (1 << r) - 1; if (r) eobrun += br.get(r);
                    break;
                }
                k += r;
                if (k > 63) return fail("AC coefficient index out of range");
                blk[kZigzag[k]] = (int16_t)(extend(br.get(s), s) * (1 << al));
                ++k;
            }
            return true;
        }
        // successive-approximation refinement of the AC band
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        auto refine = [&](int16_t &v) { if (br.get(1) && (v & p1) == 0) v = (int16_t)(v >= 0 ? v + p1 : v + m1); };
        if (eobrun == 0) {
            for (; k <= se; ++k) {
                const int rs = huff_decode(br, *ha);
                if (rs < 0) return fail("bad AC code");
                int r = rs >> 4, s = rs & 15;
                if (s) s = br.get(1) ? p1 : m1;
                else if (r != 15) { eobrun = 1 << r; if (r) eobrun += br.get(r); break; }
                do {
                    int16_t &v = blk[kZigzag[k]];
                    if (v != 0) refine(v);
                    else if (--r < 0) break;
                    ++k;
                } while (k <= se);
                if (s && k <= se) blk[kZigzag[k]] = (int16_t)s;
            }
        }
        if (eobrun > 0) {
            for (; k <= se; ++k) { int16_t &v = blk[kZigzag[k]]; if (v != 0) refine(v); }
            --eobrun;
        }
        return true;
    }

    bool decode_scan(size_t &pos) {
        const size_t len = (size_t)be16(d + pos);
        const uint8_t *s = d + pos + 2;
        if (len < 3) return fail("bad scan header");
        const int ns = s[0];
        if (ns < 1 || ns > ncomp || len != (size_t)6 + 2 * ns) return fail("bad scan header");
        int ci[3]; const HuffTable *hd[3], *ha[3];
        for (int i = 0; i < ns; ++i) {
            int k = 0; while (k < ncomp && comp[k].id != s[1 + 2 * i]) ++k;
            if (k == ncomp) return fail("scan names an unknown component");
            ci[i] = k; hd[i] = &dc[s[2 + 2 * i] >> 4 & 3]; ha[i] = &ac[s[2 + 2 * i] & 3];
        }
        const int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
        if (progressive) { if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13) return fail("bad progressive scan parameters"); }
        else if (ss != 0 || se != 63 || ah != 0 || al != 0) return fail("bad sequential scan parameters");
        for (int i = 0; i < ns; ++i) {
            if ((!progressive || (ss == 0 && ah == 0)) && !hd[i]->present) return fail("scan uses a DC table that was not defined");
            if ((!progressive || ss > 0) && !ha[i]->present) return fail("scan uses an AC table that was not defined");
            if (!have_qt[comp[ci[i]].tq]) return fail("component uses a quantisation table that was not defined");
        }
        pos += len;
        BitReader br{d, n, pos};
        for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
        eobrun = 0;
        int next_rst = 0, until_restart = restart_interval;
        auto one = [&](Component &c, int i, int bx, int by) -> bool {
            int16_t *blk = &c.coef[((size_t)by * c.bw + bx) * 64];
            return progressive ? decode_block_progressive(br, c, hd[i], ha[i], blk, ss, se, ah, al) : decode_block_baseline(br, c, *hd[i], *ha[i], blk);
        };
        if (ns == 1) {   // non-interleaved: the component's own blocks, row by row
            Component &c = comp[ci[0]];
            for (int by = 0; by < c.chh; ++by)
                for (int bx = 0; bx < c.cw; ++bx) {
                    if (restart_interval && until_restart == 0) { if (!restart(br, next_rst)) return false; until_restart = restart_interval; }
                    if (!one(c, 0, bx, by)) return false;
                    --until_restart;
                }
        } else {
            const int mcux = comp[0].bw / comp[0].h, mcuy = comp[0].bh / comp[0].v;
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    if (restart_interval && until_restart == 0) { if (!restart(br, next_rst)) return false; until_restart = restart_interval; }
                    for (int i = 0; i < ns; ++i) {
                        Component &c = comp[ci[i]];
                        for (int v = 0; v < c.v; ++v)
                            for (int h = 0; h < c.h; ++h)
                                if (!one(c, i, mx * c.h + h, my * c.v + v)) return false;
                    }
                    --until_restart;
                }
        }
        // continue after the entropy-coded segment: at the marker the reader ran into, or at the next marker in the stream
        if (br.marker) pos = br.pos - 2;
        else { size_t q = br.pos; while (q + 1 < n && !(d[q] == 0xFF && d[q + 1] != 0 && d[q + 1] != 0xFF)) ++q; pos = q; }
        return true;
    }

    void reconstruct() {
        for (int i = 0; i < ncomp; ++i) {
            Component &c = comp[i];
            c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
            parallel_rows(c.chh, (size_t)c.cw * 2048, [&](int a, int b) {
                for (int by = a; by < b; ++by)
                    for (int bx = 0; bx < c.cw; ++bx)
                        idct_islow(&c.coef[((size_t)by * c.bw + bx) * 64], qt[c.tq], &c.plane[((size_t)by * 8) * (c.bw * 8) + (size_t)bx * 8], c.bw * 8);
            });
            c.coef.clear(); c.coef.shrink_to_fit();
        }
    }

    // jdsample.c: component plane (dw x dh valid samples, row stride bw * 8) -> full resolution [height][width]
    void upsample(const Component &c, std::vector<uint8_t> &out) {
        const int W = width, H = height, hs = hmax / c.h, vs = vmax / c.v, st = c.bw * 8, dw = c.dw, dh = c.dh;
        out.assign((size_t)W * H, 0);
        const uint8_t *in = c.plane.data();
        auto row = [&](int r) { return in + (size_t)std::min(std::max(r, 0), dh - 1) * st; };   // rows above the top / below the bottom replicate the edge row
        parallel_rows(H, (size_t)W * 16, [&](int ya, int yb) {
        std::vector<uint8_t> line((size_t)dw * hs + 8);
        if (hs == 1 && vs == 1) {
            for (int y = ya; y < yb; ++y) memcpy(&out[(size_t)y * W], row(y), (size_t)W);
        } else if (hs == 2 && vs == 1) {   // h2v1_fancy_upsample
            for (int y = ya; y < yb; ++y) {
                const uint8_t *p = row(y); uint8_t *o = line.data();
                if (dw == 1) { o[0] = o[1] = p[0]; }
                else {
                    o[0] = p[0]; o[1] = (uint8_t)((p[0] * 3 + p[1] + 2) >> 2);
                    for (int x = 1; x < dw - 1; ++x) { const int v = p[x] * 3; o[2 * x] = (uint8_t)((v + p[x - 1] + 1) >> 2); o[2 * x + 1] = (uint8_t)((v + p[x + 1] + 2) >> 2); }
                    o[2 * dw - 2] = (uint8_t)((p[dw - 1] * 3 + p[dw - 2] + 1) >> 2); o[2 * dw - 1] = p[dw - 1];
                }
                memcpy(&out[(size_t)y * W], o, (size_t)W);
            }
        } else if (hs == 2 && vs == 2) {   // h2v2_fancy_upsample
            for (int y = ya; y < yb; ++y) {
                const int r = y >> 1;
                const uint8_t *p0 = row(r), *p1 = (y & 1) ? row(r + 1) : row(r - 1);
                uint8_t *o = line.data();
                if (dw == 1) { const int t = p0[0] * 3 + p1[0]; o[0] = (uint8_t)((t * 4 + 8) >> 4); o[1] = (uint8_t)((t * 4 + 7) >> 4); }
                else {
                    int thiscol = p0[0] * 3 + p1[0], nextcol = p0[1] * 3 + p1[1], lastcol;
                    o[0] = (uint8_t)((thiscol * 4 + 8) >> 4); o[1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
                    lastcol = thiscol; thiscol = nextcol;
                    for (int x = 1; x < dw - 1; ++x) {
                        nextcol = p0[x + 1] * 3 + p1[x + 1];
                        o[2 * x] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4); o[2 * x + 1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
                        lastcol = thiscol; thiscol = nextcol;
                    }
                    o[2 * dw - 2] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4); o[2 * dw - 1] = (uint8_t)((thiscol * 4 + 7) >> 4);
                }
                memcpy(&out[(size_t)y * W], o, (size_t)W);
            }
        } else if (hs == 1 && vs == 2) {   // h1v2_fancy_upsample
            for (int y = ya; y < yb; ++y) {
                const int r = y >> 1;
                const uint8_t *p0 = row(r), *p1 = (y & 1) ? row(r + 1) : row(r - 1);
                const int bias = (y & 1) ? 2 : 1;
                for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = (uint8_t)((p0[x] * 3 + p1[x] + bias) >> 2);
            }
        } else {   // int_upsample: pixel replication for every other integral ratio
            for (int y = ya; y < yb; ++y) { const uint8_t *p = row(y / vs); for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = p[x / hs]; }
        }
        });
    }

    bool run(RgbImage &img) {
        if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("not a JPEG file");
        size_t pos = 2; bool saw_scan = false;
        for (;;) {
            int m = 0;
            if (!read_tables_and_header(pos, m)) { if (saw_scan && have_sof) break; return false; }   // (a file cut after its last scan still shows what was decoded, like libjpeg with a warning)
            if (m == 0xD9) break;
            if (!have_sof) return fail("scan before the frame header");
            if (!decode_scan(pos)) return false;
            saw_scan = true;
        }
        if (!have_sof || !saw_scan) return fail("JPEG without image data");
        reconstruct();
        std::vector<uint8_t> full[3];
        for (int i = 0; i < ncomp; ++i) upsample(comp[i], full[i]);
        const size_t np = (size_t)width * height;
        std::vector<uint8_t> rgb(np * 3);
        bool ycc = ncomp == 3;
        if (ncomp == 3) {
            if (adobe) ycc = adobe_transform != 0;
            else if (!jfif && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B') ycc = false;
        }
        if (ncomp == 1) { for (size_t i = 0; i < np; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = full[0][i]; }
        else if (!ycc) { for (size_t i = 0; i < np; ++i) { rgb[3 * i] = full[0][i]; rgb[3 * i + 1] = full[1][i]; rgb[3 * i + 2] = full[2][i]; } }
        else {   // jdcolor.c build_ycc_rgb_table / ycc_rgb_convert
            int crr[256], cbb[256]; long crg[256], cbg[256];
            for (int i = 0; i < 256; ++i) {
                const long x = i - 128;
                crr[i] = (int)((91881L * x + 32768) >> 16); cbb[i] = (int)((116130L * x + 32768) >> 16);
                crg[i] = -46802L * x; cbg[i] = -22554L * x + 32768;
            }
            parallel_rows(height, (size_t)width * 16, [&](int ya, int yb) {
                for (size_t i = (size_t)ya * width; i < (size_t)yb * width; ++i) {
                    const int y = full[0][i], cb = full[1][i], cr = full[2][i];
                    rgb[3 * i] = clamp8(y + crr[cr]); rgb[3 * i + 1] = clamp8(y + (int)((cbg[cb] + crg[cr]) >> 16)); rgb[3 * i + 2] = clamp8(y + cbb[cb]);
                }
            });
        }
        // EXIF orientation (cv::imread applies it unless asked not to): 1 as stored, 2 mirrored, 3 rotated 180, 4 flipped, 5 transposed, 6 rotated 90 cw, 7 transverse, 8 rotated 270 cw
        const int W = width, H = height, o = orientation;
        const bool swap = o >= 5;
        img.w = swap ? H : W; img.h = swap ? W : H; img.px.resize(np * 3);
        for (int y = 0; y < img.h; ++y)
            for (int x = 0; x < img.w; ++x) {
                int sx, sy;
                switch (o) {
                    case 2: sx = W - 1 - x; sy = y; break;
                    case 3: sx = W - 1 - x; sy = H - 1 - y; break;
                    case 4: sx = x; sy = H - 1 - y; break;
                    case 5: sx = y; sy = x; break;
                    case 6: sx = y; sy = H - 1 - x; break;
                    case 7: sx = W - 1 - y; sy = H - 1 - x; break;
                    case 8: sx = W - 1 - y; sy = x; break;
                    default: sx = x; sy = y;
                }
                memcpy(&img.px[((size_t)y * img.w + x) * 3], &rgb[((size_t)sy * W + sx) * 3], 3);
            }
        return true;
    }
};

}  // namespace

bool decode_jpeg(const uint8_t *data, size_t n, RgbImage &out, std::string &err) {
    Decoder dec{data, n, err};
    return dec.run(out);
}

}  // namespace mg4
