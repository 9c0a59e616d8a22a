// rrt_image.cpp -- native skybox image decoder behind rrt_image_load / rrt_sky_load (include/rrt.h): the step that
// feeds the texture the render path samples (reference loadSkybox, src/main.cpp:237-266, which calls
// stbi_load(path, &w, &h, &c, 4) at :240).
//
// Contract: the RGBA8 bytes returned here are the bytes stb_image v2.30 returns for the same file with req_comp = 4,
// so a frame rendered from a file is the frame the reference renders from it.  For PNG that follows from the format
// being lossless; for JPEG the standard leaves the IDCT, the chroma upsampling filter and the YCbCr->RGB rounding to
// the decoder, so those three stages use the arithmetic stb_image documents for itself:
//   * IDCT: the jidctint "ISLOW" factorisation with 12-bit constants, column pass keeping 2 extra bits (+512 >> 10), row
//     pass folding the +128 level shift into the rounding term (+65536 + (128 << 17)) >> 17, then clamp;
//   * upsampling: JFIF-centred triangle filters (3:1 vertically, 3:1 horizontally, 9:3:3:1 for 2x2), nearest otherwise;
//   * colour: 20-bit fixed point, Y << 20 + 2^19, coefficients round(c * 4096) << 8, the Cb term of G masked to its high
//     16 bits (which is what makes stb's own scalar and SIMD kernels agree).
// tests/test_image_decoder.py checks byte equality against stb_image itself (compiled from the reference tree as a
// test-only checker library) on the reference's two assets and on generated files.
//
// Scope: JPEG baseline / extended sequential Huffman, 8 bit, 1 or 3 components, any sampling factors, restart intervals,
// interleaved and single-component scans; PNG 8 bits per channel (grey, grey+alpha, RGB, RGBA) and 1/2/4/8-bit
// palette or grey, tRNS, non-interlaced.  Progressive / arithmetic / 12-bit / CMYK JPEG and 16-bit or Adam7 PNG are
// reported as RRT_ERR_UNSUPPORTED rather than decoded differently from the reference.
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rrt.h"

namespace {

thread_local std::string g_img_err;
int img_fail(int code, const char* what) {
    g_img_err = what;
    return code;
}

struct Bytes {
    const uint8_t* p;
    size_t n, pos;
    bool eof() const { return pos >= n; }
    int u8() { return pos < n ? p[pos++] : 0; }
    int be16() { int a = u8(); return (a << 8) | u8(); }
    uint32_t be32() { uint32_t a = (uint32_t)be16(); return (a << 16) | (uint32_t)be16(); }
};

// =====================================================================================================
// JPEG
// =====================================================================================================
const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool present = false;
    uint8_t vals[256];
    int mincode[17], maxcode[18], valptr[17];   // ITU T.81 F.2.2.3
    int16_t look[512];                          // 9-bit prefix -> (len << 8 | symbol), or -1
    bool build(const uint8_t counts[16], const uint8_t* symbols, int nsym) {
        std::memcpy(vals, symbols, (size_t)nsym);
        int code = 0, k = 0;
        for (int i = 0; i < 512; ++i) look[i] = -1;
        for (int len = 1; len <= 16; ++len) {
            valptr[len] = k;
            mincode[len] = code;
            for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
                if (len <= 9) {
                    const int lo = code << (9 - len), hi = lo + (1 << (9 - len));
                    for (int c = lo; c < hi; ++c) look[c] = (int16_t)((len << 8) | vals[k]);
                }
            }
            if (code > (1 << len)) return false;
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
        return k == nsym;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int w2 = 0, h2 = 0;   // plane size, padded to whole MCUs
    int x = 0, y = 0;     // real size of this component
    int dc_pred = 0;
    std::vector<uint8_t> plane;
};

struct BitReader {
    Bytes* src;
    uint64_t acc = 0;
    int nbits = 0;
    int marker = 0;     // marker met inside the entropy-coded segment (0 = none); zeros are fed after it
    void reset() { acc = 0; nbits = 0; marker = 0; }
    void fill() {
        while (nbits <= 48) {
            int b = 0;
            if (!marker && !src->eof()) {
                b = src->u8();
                if (b == 0xFF) {
                    int c = src->u8();
                    while (c == 0xFF) c = src->u8();      // fill bytes
                    if (c != 0) { marker = c; b = 0; }
                }
            }
            acc |= (uint64_t)b << (56 - nbits);
            nbits += 8;
        }
    }
    int peek(int n) { return (int)(acc >> (64 - n)); }
    void drop(int n) { acc <<= n; nbits -= n; }
    int bits(int n) {
        if (n == 0) return 0;
        if (nbits < n) fill();
        int v = peek(n);
        drop(n);
        return v;
    }
    int symbol(const Huff& h) {
        if (nbits < 16) fill();
        const int e = h.look[peek(9)];
        if (e >= 0) { drop(e >> 8); return e & 255; }
        int code = peek(9), len = 9;
        for (;;) {
            ++len;
            if (len > 16) return -1;
            code = peek(len);
            if (h.maxcode[len] >= 0 && code <= h.maxcode[len]) break;
        }
        drop(len);
        return h.vals[h.valptr[len] + code - h.mincode[len]];
    }
    int extend(int s) {   // T.81 F.2.2.1 RECEIVE + EXTEND
        const int v = bits(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
};

inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// One 1-D pass of the ISLOW butterfly on 8 inputs; writes the even part sums e[0..3] and the odd part o[0..3] such
// that out[k] = e[k] + o[3-k], out[7-k] = e[k] - o[3-k] (before the pass-specific rounding shift).
inline int fx(double c) { return (int)(c * 4096 + 0.5); }
inline void islow_1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7, int e[4], int o[4]) {
    static const int c0541 = fx(0.5411961f), c1847 = fx(-1.847759065f), c0765 = fx(0.765366865f), c1175 = fx(1.175875602f),
                     c0298 = fx(0.298631336f), c2053 = fx(2.053119869f), c3072 = fx(3.072711026f), c1501 = fx(1.501321110f),
                     c0899 = fx(-0.899976223f), c2562 = fx(-2.562915447f), c1961 = fx(-1.961570560f), c0390 = fx(-0.390180644f);
    const int z = (s2 + s6) * c0541;
    const int a2 = z + s6 * c1847, a3 = z + s2 * c0765;
    const int a0 = (s0 + s4) * 4096, a1 = (s0 - s4) * 4096;
    e[0] = a0 + a3; e[3] = a0 - a3; e[1] = a1 + a2; e[2] = a1 - a2;
    const int q3 = s7 + s3, q4 = s5 + s1, q1 = s7 + s1, q2 = s5 + s3;
    const int z5 = (q3 + q4) * c1175;
    const int m1 = z5 + q1 * c0899, m2 = z5 + q2 * c2562, m3 = q3 * c1961, m4 = q4 * c0390;
    o[3] = s1 * c1501 + (m1 + m4);
    o[2] = s3 * c3072 + (m2 + m3);
    o[1] = s5 * c2053 + (m2 + m4);
    o[0] = s7 * c0298 + (m1 + m3);
}
void idct_8x8(uint8_t* out, int stride, const int16_t d[64]) {
    int tmp[64];
    for (int c = 0; c < 8; ++c) {
        int e[4], o[4];
        islow_1d(d[c], d[8 + c], d[16 + c], d[24 + c], d[32 + c], d[40 + c], d[48 + c], d[56 + c], e, o);
        for (int k = 0; k < 4; ++k) {
            tmp[8 * k + c] = (e[k] + 512 + o[3 - k]) >> 10;
            tmp[8 * (7 - k) + c] = (e[k] + 512 - o[3 - k]) >> 10;
        }
    }
    const int bias = 65536 + (128 << 17);
    for (int r = 0; r < 8; ++r, out += stride) {
        const int* v = tmp + 8 * r;
        int e[4], o[4];
        islow_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], e, o);
        for (int k = 0; k < 4; ++k) {
            out[k] = clamp8((e[k] + bias + o[3 - k]) >> 17);
            out[7 - k] = clamp8((e[k] + bias - o[3 - k]) >> 17);
        }
    }
}

struct Jpeg {
    Bytes in;
    int w = 0, h = 0, ncomp = 0;
    int hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0;
    int restart_interval = 0;
    bool jfif = false;
    int adobe_transform = -1;
    int rgb_ids = 0;
    uint16_t quant[4][64];
    bool quant_set[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Component comp[4];
    bool have_frame = false;

    int block(BitReader& br, Component& c, int16_t data[64]) {
        const Huff &hd = dc[c.td], &ha = ac[c.ta];
        if (!hd.present || !ha.present || !quant_set[c.tq]) return img_fail(RRT_ERR_IO, "JPEG: scan uses a table that was not defined");
        std::memset(data, 0, 64 * sizeof(int16_t));
        const int t = br.symbol(hd);
        if (t < 0 || t > 15) return img_fail(RRT_ERR_IO, "JPEG: bad DC Huffman code");
        c.dc_pred += t ? br.extend(t) : 0;
        data[0] = (int16_t)(c.dc_pred * quant[c.tq][0]);
        for (int k = 1; k < 64;) {
            const int rs = br.symbol(ha);
            if (rs < 0) return img_fail(RRT_ERR_IO, "JPEG: bad AC Huffman code");
            const int run = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (rs != 0xF0) break;   // end of block
                k += 16;
                continue;
            }
            k += run;
            if (k > 63) return img_fail(RRT_ERR_IO, "JPEG: coefficient index out of range");
            const int zz = kZigzag[k++];
            data[zz] = (int16_t)(br.extend(s) * quant[c.tq][zz]);
        }
        return RRT_OK;
    }
};

int jpeg_read_tables_and_frame(Jpeg& J, int marker, int len) {
    Bytes& in = J.in;
    const size_t end = in.pos + (size_t)len;
    if (end > in.n) return img_fail(RRT_ERR_IO, "JPEG: truncated segment");
    switch (marker) {
        case 0xDB:   // DQT
            while (in.pos < end) {
                const int pq = in.u8(), prec = pq >> 4, t = pq & 15;
                if (prec > 1 || t > 3) return img_fail(RRT_ERR_IO, "JPEG: bad DQT");
                for (int i = 0; i < 64; ++i) J.quant[t][kZigzag[i]] = (uint16_t)(prec ? in.be16() : in.u8());
                J.quant_set[t] = true;
            }
            break;
        case 0xC4:   // DHT
            while (in.pos < end) {
                const int tc = in.u8(), cls = tc >> 4, t = tc & 15;
                uint8_t counts[16];
                int n = 0;
                for (int i = 0; i < 16; ++i) { counts[i] = (uint8_t)in.u8(); n += counts[i]; }
                if (cls > 1 || t > 3 || n > 256 || in.pos + (size_t)n > end) return img_fail(RRT_ERR_IO, "JPEG: bad DHT");
                if (!(cls ? J.ac[t] : J.dc[t]).build(counts, in.p + in.pos, n)) return img_fail(RRT_ERR_IO, "JPEG: bad Huffman code lengths");
                in.pos += (size_t)n;
            }
            break;
        case 0xDD:   // DRI
            J.restart_interval = in.be16();
            break;
        case 0xE0:   // APP0
            if (len >= 5 && std::memcmp(in.p + in.pos, "JFIF\0", 5) == 0) J.jfif = true;
            break;
        case 0xEE:   // APP14
            if (len >= 12 && std::memcmp(in.p + in.pos, "Adobe\0", 6) == 0) J.adobe_transform = in.p[in.pos + 11];
            break;
        case 0xC0: case 0xC1: {   // SOF0 / SOF1
            if (J.have_frame) return img_fail(RRT_ERR_IO, "JPEG: two frame headers");
            if (in.u8() != 8) return img_fail(RRT_ERR_UNSUPPORTED, "JPEG: only 8-bit precision is supported");
            J.h = in.be16();
            J.w = in.be16();
            J.ncomp = in.u8();
            if (J.w <= 0 || J.h <= 0) return img_fail(RRT_ERR_UNSUPPORTED, "JPEG: zero size (DNL-defined height is not supported)");
            if (J.ncomp == 4) return img_fail(RRT_ERR_UNSUPPORTED, "JPEG: CMYK / YCCK is not supported");
            if (J.ncomp != 1 && J.ncomp != 3) return img_fail(RRT_ERR_IO, "JPEG: bad component count");
            if (len != 6 + 3 * J.ncomp) return img_fail(RRT_ERR_IO, "JPEG: bad SOF length");
            static const char rgb[3] = {'R', 'G', 'B'};
            for (int i = 0; i < J.ncomp; ++i) {
                Component& c = J.comp[i];
                c.id = in.u8();
                if (J.ncomp == 3 && c.id == rgb[i]) ++J.rgb_ids;
                const int hv = in.u8();
                c.h = hv >> 4; c.v = hv & 15; c.tq = in.u8();
                if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return img_fail(RRT_ERR_IO, "JPEG: bad sampling factors");
                if (c.h > J.hmax) J.hmax = c.h;
                if (c.v > J.vmax) J.vmax = c.v;
            }
            for (int i = 0; i < J.ncomp; ++i)
                if (J.hmax % J.comp[i].h || J.vmax % J.comp[i].v) return img_fail(RRT_ERR_IO, "JPEG: fractional sampling ratio");
            J.mcus_x = (J.w + 8 * J.hmax - 1) / (8 * J.hmax);
            J.mcus_y = (J.h + 8 * J.vmax - 1) / (8 * J.vmax);
            for (int i = 0; i < J.ncomp; ++i) {
                Component& c = J.comp[i];
                c.x = (J.w * c.h + J.hmax - 1) / J.hmax;
                c.y = (J.h * c.v + J.vmax - 1) / J.vmax;
                c.w2 = J.mcus_x * c.h * 8;
                c.h2 = J.mcus_y * c.v * 8;
                c.plane.assign((size_t)c.w2 * c.h2, 0);
            }
            J.have_frame = true;
            break;
        }
        default:
            break;   // other APPn / COM: skipped
    }
    in.pos = end;
    return RRT_OK;
}

int jpeg_decode_scan(Jpeg& J) {
    Bytes& in = J.in;
    const int len = in.be16();
    const int ns = in.u8();
    if (!J.have_frame) return img_fail(RRT_ERR_IO, "JPEG: scan before frame header");
    if (ns < 1 || ns > J.ncomp || len != 6 + 2 * ns) return img_fail(RRT_ERR_IO, "JPEG: bad SOS");
    int order[4];
    for (int i = 0; i < ns; ++i) {
        const int id = in.u8(), tt = in.u8();
        int which = -1;
        for (int k = 0; k < J.ncomp; ++k)
            if (J.comp[k].id == id) which = k;
        if (which < 0) return img_fail(RRT_ERR_IO, "JPEG: scan names an unknown component");
        J.comp[which].td = tt >> 4;
        J.comp[which].ta = tt & 15;
        if (J.comp[which].td > 3 || J.comp[which].ta > 3) return img_fail(RRT_ERR_IO, "JPEG: bad table selector");
        order[i] = which;
    }
    const int ss = in.u8(), se = in.u8(), ahal = in.u8();
    if (ss != 0 || se != 63 || ahal != 0) return img_fail(RRT_ERR_IO, "JPEG: spectral selection in a sequential scan");

    BitReader br;
    br.src = &in;
    for (int k = 0; k < J.ncomp; ++k) J.comp[k].dc_pred = 0;
    int16_t data[64];
    int todo = J.restart_interval ? J.restart_interval : 0x7fffffff;
    // after `todo` MCUs: byte-align, expect RSTn, reset predictors; anything else ends the scan
    auto restart = [&]() -> bool {
        if (--todo > 0) return true;
        br.fill();
        if (br.marker < 0xD0 || br.marker > 0xD7) return false;
        br.reset();
        for (int k = 0; k < J.ncomp; ++k) J.comp[k].dc_pred = 0;
        todo = J.restart_interval ? J.restart_interval : 0x7fffffff;
        return true;
    };
    bool more = true;
    if (ns == 1) {   // single-component scan: the component's own block raster
        Component& c = J.comp[order[0]];
        const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
        for (int by = 0; by < bh && more; ++by)
            for (int bx = 0; bx < bw && more; ++bx) {
                const int rc = J.block(br, c, data);
                if (rc) return rc;
                idct_8x8(c.plane.data() + (size_t)by * 8 * c.w2 + bx * 8, c.w2, data);
                more = restart();
            }
    } else {
        for (int my = 0; my < J.mcus_y && more; ++my)
            for (int mx = 0; mx < J.mcus_x && more; ++mx) {
                for (int i = 0; i < ns; ++i) {
                    Component& c = J.comp[order[i]];
                    for (int v = 0; v < c.v; ++v)
                        for (int hh = 0; hh < c.h; ++hh) {
                            const int rc = J.block(br, c, data);
                            if (rc) return rc;
                            idct_8x8(c.plane.data() + (size_t)(my * c.v + v) * 8 * c.w2 + (mx * c.h + hh) * 8, c.w2, data);
                        }
                }
                more = restart();
            }
    }
    // hand a marker the bit reader ran into back to the segment loop
    br.fill();
    if (br.marker) return 0x100 | br.marker;   // its bytes are consumed: reported through the return value
    return RRT_OK;
}

// chroma / luma row at full width, stb_image's JFIF-centred filters
const uint8_t* upsample_row(uint8_t* out, const uint8_t* near_, const uint8_t* far_, int w, int hs, int vs) {
    if (hs == 1 && vs == 1) return near_;
    if (hs == 1 && vs == 2) {
        for (int i = 0; i < w; ++i) out[i] = (uint8_t)((3 * near_[i] + far_[i] + 2) >> 2);
        return out;
    }
    if (hs == 2 && vs == 1) {
        if (w == 1) { out[0] = out[1] = near_[0]; return out; }
        out[0] = near_[0];
        out[1] = (uint8_t)((near_[0] * 3 + near_[1] + 2) >> 2);
        int i = 1;
        for (; i < w - 1; ++i) {
            const int n = 3 * near_[i] + 2;
            out[2 * i] = (uint8_t)((n + near_[i - 1]) >> 2);
            out[2 * i + 1] = (uint8_t)((n + near_[i + 1]) >> 2);
        }
        out[2 * i] = (uint8_t)((near_[w - 2] * 3 + near_[w - 1] + 2) >> 2);
        out[2 * i + 1] = near_[w - 1];
        return out;
    }
    if (hs == 2 && vs == 2) {
        if (w == 1) { out[0] = out[1] = (uint8_t)((3 * near_[0] + far_[0] + 2) >> 2); return out; }
        int t1 = 3 * near_[0] + far_[0];
        out[0] = (uint8_t)((t1 + 2) >> 2);
        for (int i = 1; i < w; ++i) {
            const int t0 = t1;
            t1 = 3 * near_[i] + far_[i];
            out[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            out[2 * i] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        out[2 * w - 1] = (uint8_t)((t1 + 2) >> 2);
        return out;
    }
    for (int i = 0; i < w; ++i)
        for (int j = 0; j < hs; ++j) out[i * hs + j] = near_[i];
    return out;
}

inline int cfix(float c) { return ((int)(c * 4096.0f + 0.5f)) << 8; }
void ycc_row_to_rgba(uint8_t* out, const uint8_t* y, const uint8_t* cb, const uint8_t* cr, int n) {
    const int kr = cfix(1.40200f), kg_r = -cfix(0.71414f), kg_b = -cfix(0.34414f), kb = cfix(1.77200f);
    for (int i = 0; i < n; ++i, out += 4) {
        const int yf = (y[i] << 20) + (1 << 19);
        const int r_ = cr[i] - 128, b_ = cb[i] - 128;
        int r = yf + r_ * kr;
        int g = yf + r_ * kg_r + (int)((unsigned)(b_ * kg_b) & 0xffff0000u);
        int b = yf + b_ * kb;
        r >>= 20; g >>= 20; b >>= 20;
        out[0] = clamp8(r); out[1] = clamp8(g); out[2] = clamp8(b); out[3] = 255;
    }
}

int decode_jpeg(const uint8_t* buf, size_t n, std::vector<uint8_t>& rgba, int& w, int& h) {
    Jpeg J;
    J.in = Bytes{buf, n, 0};
    Bytes& in = J.in;
    if (in.u8() != 0xFF || in.u8() != 0xD8) return img_fail(RRT_ERR_IO, "JPEG: no SOI");
    int pending = 0;   // marker already consumed by the entropy decoder
    bool scanned = false;
    for (;;) {
        int m = pending;
        pending = 0;
        if (!m) {
            // next marker; stray bytes between segments are skipped like stb_image does after a scan
            int b = in.eof() ? -1 : in.u8();
            while (b >= 0 && b != 0xFF) b = in.eof() ? -1 : in.u8();
            if (b < 0) break;
            do { m = in.eof() ? 0xD9 : in.u8(); } while (m == 0xFF);
            if (m == 0) continue;
        }
        if (m == 0xD9) break;                              // EOI
        if (m >= 0xD0 && m <= 0xD7) continue;              // stray RSTn
        if (m == 0xC2) return img_fail(RRT_ERR_UNSUPPORTED, "JPEG: progressive files are not supported");
        if (m == 0xC9 || m == 0xCA || m == 0xCB || m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xCD && m <= 0xCF))
            return img_fail(RRT_ERR_UNSUPPORTED, "JPEG: arithmetic / lossless / hierarchical coding is not supported");
        if (m == 0xDA) {
            const int rc = jpeg_decode_scan(J);
            if (rc < 0) return rc;
            if (rc & 0x100) pending = rc & 0xFF;
            scanned = true;
            continue;
        }
        const int len = in.be16();
        if (len < 2) return img_fail(RRT_ERR_IO, "JPEG: bad segment length");
        const int rc = jpeg_read_tables_and_frame(J, m, len - 2);
        if (rc) return rc;
    }
    if (!J.have_frame || !scanned) return img_fail(RRT_ERR_IO, "JPEG: no image data");

    w = J.w; h = J.h;
    rgba.assign((size_t)w * h * 4, 255);
    const bool is_rgb = J.ncomp == 3 && (J.rgb_ids == 3 || (J.adobe_transform == 0 && !J.jfif));
    struct Up { int hs, vs, ystep, ypos, wlo; const uint8_t *l0, *l1; std::vector<uint8_t> buf; } up[3];
    for (int k = 0; k < J.ncomp; ++k) {
        Component& c = J.comp[k];
        up[k].hs = J.hmax / c.h; up[k].vs = J.vmax / c.v;
        up[k].ystep = up[k].vs >> 1; up[k].ypos = 0;
        up[k].wlo = (w + up[k].hs - 1) / up[k].hs;
        up[k].l0 = up[k].l1 = c.plane.data();
        up[k].buf.assign((size_t)w + 8, 0);
    }
    for (int j = 0; j < h; ++j) {
        const uint8_t* row[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < J.ncomp; ++k) {
            Up& u = up[k];
            const bool bot = u.ystep >= (u.vs >> 1);
            row[k] = upsample_row(u.buf.data(), bot ? u.l1 : u.l0, bot ? u.l0 : u.l1, u.wlo, u.hs, u.vs);
            if (++u.ystep >= u.vs) {
                u.ystep = 0;
                u.l0 = u.l1;
                if (++u.ypos < J.comp[k].y) u.l1 += J.comp[k].w2;
            }
        }
        uint8_t* out = rgba.data() + (size_t)j * w * 4;
        if (J.ncomp == 1) {
            for (int i = 0; i < w; ++i, out += 4) { out[0] = out[1] = out[2] = row[0][i]; out[3] = 255; }
        } else if (is_rgb) {
            for (int i = 0; i < w; ++i, out += 4) { out[0] = row[0][i]; out[1] = row[1][i]; out[2] = row[2][i]; out[3] = 255; }
        } else {
            ycc_row_to_rgba(out, row[0], row[1], row[2], w);
        }
    }
    return RRT_OK;
}

// =====================================================================================================
// PNG
// =====================================================================================================
struct Inflate {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t acc = 0;
    int nb = 0;
    std::vector<uint8_t>& out;
    Inflate(const uint8_t* p_, size_t n_, std::vector<uint8_t>& o) : p(p_), n(n_), out(o) {}
    int bits(int k) {
        while (nb < k) { acc |= (uint32_t)(pos < n ? p[pos++] : 0) << nb; nb += 8; }
        const int v = (int)(acc & ((1u << k) - 1));
        acc >>= k; nb -= k;
        return v;
    }
    struct Table { uint16_t count[16]; uint16_t sym[288]; };
    static bool build(Table& t, const uint8_t* len, int n) {
        std::memset(t.count, 0, sizeof(t.count));
        for (int i = 0; i < n; ++i) t.count[len[i]]++;
        t.count[0] = 0;
        int offs[16], left = 1;
        for (int l = 1; l < 16; ++l) { left = (left << 1) - t.count[l]; if (left < 0) return false; }
        offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + t.count[l];
        for (int i = 0; i < n; ++i) if (len[i]) t.sym[offs[len[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(const Table& t) {   // canonical, bit by bit (codes are packed LSB first, MSB of the code first)
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; ++l) {
            code |= bits(1);
            const int c = t.count[l];
            if (code - c < first) return t.sym[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
    bool run() {
        static const uint16_t lbase[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
        static const uint8_t lext[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
        static const uint16_t dbase[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
        static const uint8_t dext[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
        int last;
        do {
            last = bits(1);
            const int type = bits(2);
            if (type == 0) {
                acc = 0; nb = 0;
                if (pos + 4 > n) return false;
                const int len = p[pos] | (p[pos + 1] << 8), nlen = p[pos + 2] | (p[pos + 3] << 8);
                pos += 4;
                if ((len ^ 0xffff) != nlen || pos + (size_t)len > n) return false;
                out.insert(out.end(), p + pos, p + pos + len);
                pos += (size_t)len;
                continue;
            }
            if (type == 3) return false;
            Table tl, td;
            uint8_t lens[320];
            if (type == 1) {
                int i = 0;
                for (; i < 144; ++i) lens[i] = 8;
                for (; i < 256; ++i) lens[i] = 9;
                for (; i < 280; ++i) lens[i] = 7;
                for (; i < 288; ++i) lens[i] = 8;
                build(tl, lens, 288);
                for (i = 0; i < 30; ++i) lens[i] = 5;
                build(td, lens, 30);
            } else {
                static const uint8_t order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
                const int nlen = bits(5) + 257, ndist = bits(5) + 1, ncode = bits(4) + 4;
                if (nlen > 286 || ndist > 30) return false;
                uint8_t cl[19] = {0};
                for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)bits(3);
                Table tc;
                if (!build(tc, cl, 19)) return false;
                int i = 0;
                while (i < nlen + ndist) {
                    const int s = decode(tc);
                    if (s < 0) return false;
                    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                    int rep, val = 0;
                    if (s == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + bits(2); }
                    else if (s == 17) rep = 3 + bits(3);
                    else rep = 11 + bits(7);
                    if (i + rep > nlen + ndist) return false;
                    while (rep--) lens[i++] = (uint8_t)val;
                }
                if (!build(tl, lens, nlen) || !build(td, lens + nlen, ndist)) return false;
            }
            for (;;) {
                const int s = decode(tl);
                if (s < 0) return false;
                if (s < 256) { out.push_back((uint8_t)s); continue; }
                if (s == 256) break;
                if (s > 285) return false;
                const int len = lbase[s - 257] + bits(lext[s - 257]);
                const int ds = decode(td);
                if (ds < 0 || ds > 29) return false;
                const size_t dist = (size_t)dbase[ds] + (size_t)bits(dext[ds]);
                if (dist > out.size()) return false;
                const size_t from = out.size() - dist;
                for (int k = 0; k < len; ++k) out.push_back(out[from + (size_t)k]);
            }
        } while (!last);
        return true;
    }
};

inline int paeth(int a, int b, int c) {
    const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

int decode_png(const uint8_t* buf, size_t n, std::vector<uint8_t>& rgba, int& w, int& h) {
    Bytes in{buf, n, 8};
    int depth = 0, ctype = 0, interlace = 0;
    bool have_hdr = false, have_trns = false;
    uint8_t pal[256][4];
    int npal = 0;
    uint8_t key[3] = {0, 0, 0};
    std::vector<uint8_t> z;
    for (bool done = false; !done;) {
        if (in.pos + 8 > in.n) return img_fail(RRT_ERR_IO, "PNG: truncated");
        const uint32_t len = in.be32(), type = in.be32();
        if (in.pos + (size_t)len + 4 > in.n) return img_fail(RRT_ERR_IO, "PNG: truncated chunk");
        const uint8_t* d = in.p + in.pos;
        switch (type) {
            case 0x49484452:   // IHDR
                if (len != 13) return img_fail(RRT_ERR_IO, "PNG: bad IHDR");
                w = (int)((uint32_t)d[0] << 24 | d[1] << 16 | d[2] << 8 | d[3]);
                h = (int)((uint32_t)d[4] << 24 | d[5] << 16 | d[6] << 8 | d[7]);
                depth = d[8]; ctype = d[9]; interlace = d[12];
                if (w <= 0 || h <= 0 || w > (1 << 24) || h > (1 << 24)) return img_fail(RRT_ERR_IO, "PNG: bad size");
                if (d[10] || d[11]) return img_fail(RRT_ERR_IO, "PNG: bad compression / filter method");
                if (depth == 16) return img_fail(RRT_ERR_UNSUPPORTED, "PNG: 16-bit channels are not supported");
                if (interlace) return img_fail(RRT_ERR_UNSUPPORTED, "PNG: Adam7 interlacing is not supported");
                if (!(ctype == 0 || ctype == 2 || ctype == 3 || ctype == 4 || ctype == 6)) return img_fail(RRT_ERR_IO, "PNG: bad colour type");
                if (!(depth == 8 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))))
                    return img_fail(RRT_ERR_IO, "PNG: bad bit depth");
                have_hdr = true;
                break;
            case 0x504c5445:   // PLTE
                if (len > 768 || len % 3) return img_fail(RRT_ERR_IO, "PNG: bad PLTE");
                npal = (int)len / 3;
                for (int i = 0; i < npal; ++i) { pal[i][0] = d[3 * i]; pal[i][1] = d[3 * i + 1]; pal[i][2] = d[3 * i + 2]; pal[i][3] = 255; }
                break;
            case 0x74524e53:   // tRNS
                if (!have_hdr) return img_fail(RRT_ERR_IO, "PNG: tRNS before IHDR");
                if (ctype == 3) {
                    if ((int)len > npal) return img_fail(RRT_ERR_IO, "PNG: bad tRNS");
                    for (uint32_t i = 0; i < len; ++i) pal[i][3] = d[i];
                    have_trns = true;
                } else if (ctype == 0 || ctype == 2) {
                    const int nc = ctype == 0 ? 1 : 3;
                    if ((int)len != 2 * nc) return img_fail(RRT_ERR_IO, "PNG: bad tRNS");
                    static const int scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
                    for (int k = 0; k < nc; ++k) key[k] = (uint8_t)(d[2 * k + 1] * scale[depth]);
                    have_trns = true;
                }
                break;
            case 0x49444154:   // IDAT
                z.insert(z.end(), d, d + len);
                break;
            case 0x49454e44:   // IEND
                done = true;
                break;
            default:
                break;
        }
        in.pos += (size_t)len + 4;
    }
    if (!have_hdr || z.size() < 2) return img_fail(RRT_ERR_IO, "PNG: no image data");
    if ((z[0] & 15) != 8 || ((z[0] << 8) | z[1]) % 31 || (z[1] & 32)) return img_fail(RRT_ERR_IO, "PNG: bad zlib header");
    std::vector<uint8_t> raw;
    const int chans = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
    const size_t stride = ((size_t)w * chans * depth + 7) / 8;
    raw.reserve((stride + 1) * (size_t)h);
    Inflate inf(z.data() + 2, z.size() - 2, raw);
    if (!inf.run() || raw.size() < (stride + 1) * (size_t)h) return img_fail(RRT_ERR_IO, "PNG: corrupt deflate stream");

    const int bpp = depth < 8 ? 1 : chans;   // filter unit in bytes
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    rgba.assign((size_t)w * h * 4, 255);
    static const int gscale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    for (int y = 0; y < h; ++y) {
        const uint8_t* src = raw.data() + (size_t)y * (stride + 1);
        const int f = src[0];
        ++src;
        if (f > 4) return img_fail(RRT_ERR_IO, "PNG: bad filter type");
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
            int v = src[i];
            switch (f) {
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: break;
            }
            cur[i] = (uint8_t)v;
        }
        uint8_t* out = rgba.data() + (size_t)y * w * 4;
        for (int x = 0; x < w; ++x, out += 4) {
            if (depth < 8) {
                const int per = 8 / depth, byte = cur[(size_t)x / per], sh = 8 - depth * (x % per + 1);
                const int v = (byte >> sh) & ((1 << depth) - 1);
                if (ctype == 3) {
                    if (v >= npal) return img_fail(RRT_ERR_IO, "PNG: palette index out of range");
                    out[0] = pal[v][0]; out[1] = pal[v][1]; out[2] = pal[v][2]; out[3] = have_trns ? pal[v][3] : 255;
                } else {
                    const uint8_t g = (uint8_t)(v * gscale[depth]);
                    out[0] = out[1] = out[2] = g;
                    out[3] = (have_trns && g == key[0]) ? 0 : 255;
                }
                continue;
            }
            const uint8_t* px = cur.data() + (size_t)x * chans;
            switch (ctype) {
                case 0: out[0] = out[1] = out[2] = px[0]; out[3] = (have_trns && px[0] == key[0]) ? 0 : 255; break;
                case 2: out[0] = px[0]; out[1] = px[1]; out[2] = px[2];
                        out[3] = (have_trns && px[0] == key[0] && px[1] == key[1] && px[2] == key[2]) ? 0 : 255; break;
                case 3: if (px[0] >= npal) return img_fail(RRT_ERR_IO, "PNG: palette index out of range");
                        out[0] = pal[px[0]][0]; out[1] = pal[px[0]][1]; out[2] = pal[px[0]][2]; out[3] = have_trns ? pal[px[0]][3] : 255; break;
                case 4: out[0] = out[1] = out[2] = px[0]; out[3] = px[1]; break;
                default: out[0] = px[0]; out[1] = px[1]; out[2] = px[2]; out[3] = px[3]; break;
            }
        }
        prev.swap(cur);
    }
    return RRT_OK;
}

}  // namespace

extern "C" {

const char* rrt_image_last_error(void) { return g_img_err.c_str(); }

int rrt_image_decode(const uint8_t* data, size_t size, uint8_t** rgba_out, int* w_out, int* h_out) {
    if (!data || !rgba_out || !w_out || !h_out) return img_fail(RRT_ERR_BAD_ARG, "rrt_image_decode: bad argument");
    *rgba_out = nullptr;
    std::vector<uint8_t> px;
    int w = 0, h = 0, rc;
    static const uint8_t png_sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (size >= 8 && std::memcmp(data, png_sig, 8) == 0) rc = decode_png(data, size, px, w, h);
    else if (size >= 3 && data[0] == 0xFF && data[1] == 0xD8) rc = decode_jpeg(data, size, px, w, h);
    else rc = img_fail(RRT_ERR_UNSUPPORTED, "image is neither PNG nor JPEG");
    if (rc != RRT_OK) return rc;
    uint8_t* o = (uint8_t*)std::malloc(px.size());
    if (!o) return img_fail(RRT_ERR_NOMEM, "out of host memory");
    std::memcpy(o, px.data(), px.size());
    *rgba_out = o; *w_out = w; *h_out = h;
    return RRT_OK;
}

int rrt_image_load(const char* path, uint8_t** rgba_out, int* w_out, int* h_out) {
    if (!path) return img_fail(RRT_ERR_BAD_ARG, "rrt_image_load: path is NULL");
    FILE* f = std::fopen(path, "rb");
    if (!f) return img_fail(RRT_ERR_IO, "rrt_image_load: cannot open file");
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t got;
    while ((got = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    std::fclose(f);
    return rrt_image_decode(buf.data(), buf.size(), rgba_out, w_out, h_out);
}

void rrt_image_free(uint8_t* rgba) { std::free(rgba); }

int rrt_sky_load(rrt_context* ctx, const char* path, rrt_sky** out) {
    if (!ctx || !out) return RRT_ERR_BAD_ARG;
    uint8_t* px = nullptr;
    int w = 0, h = 0;
    int rc = rrt_image_load(path, &px, &w, &h);
    if (rc != RRT_OK) return rc;
    rc = rrt_sky_create(ctx, px, w, h, out);
    rrt_image_free(px);
    return rc;
}

}  // extern "C"
