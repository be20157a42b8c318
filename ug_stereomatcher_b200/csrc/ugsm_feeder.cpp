// ugsm_feeder.cpp -- real-image feeder: image list -> TIFF decode -> ring of page-locked pair buffers
// (SURVEY.md 8f row 4).  Replaces the reference's camera simulator for batch runs:
//   * Settings::readStringList / Settings::nextImage (include/ug_stereomatcher/settings.h:15-68): an OpenCV
//     FileStorage XML whose first top-level node is a sequence of paths, consumed two at a time (left, right);
//     past the end of the list the reader steps back to the next-to-last entry, i.e. it repeats the last pair;
//   * RHcam_node::captureImage (src/utils/publish_images.cpp:128-211): cv::imread(path, 1) -> 8-bit, 3 channels,
//     B,G,R byte order (published under the label "rgb8" without a swap, so the matcher sees B,G,R bytes).
// The decoder below covers what cv::imread(.., 1) does for the TIFF variants cameras and OpenCV itself write:
// classic little/big-endian TIFF, strips or tiles, chunky or planar, 8-bit grey / RGB / palette, 16-bit grey
// (>> 8) and RGB ((v * 255 + 32767) / 65535), extra samples ignored unless they are alpha, compression none / LZW /
// PackBits / Deflate, horizontal predictor.  It is pinned against cv2.imread (the same OpenCV decoder) in
// tests/test_feeder.py and through the committed fixtures under tests/golden/tiff/.
//
// Decode threads run ahead of the consumers; every slot holds one pair in memory from ugsm_host_alloc, so
// ugsm_match_* copies it with plain asynchronous DMA.  Host-only code: no kernel, no device state.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#if defined(__SSSE3__)
#include <tmmintrin.h>
#endif

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ugsm.h"

namespace {

thread_local char g_feeder_error[512] = "";

int ffail(int code, const char* fmt, const char* a = "", long b = 0) {
    snprintf(g_feeder_error, sizeof g_feeder_error, fmt, a, b);
    return code;
}

// ------------------------------------------------------------------------------------------------ TIFF
struct TiffInfo {
    bool be = false;
    uint32_t width = 0, height = 0, spp = 1, bps = 8, compression = 1, photometric = 1, planar = 1, predictor = 1;
    uint32_t rows_per_strip = 0xffffffffu, tile_w = 0, tile_h = 0, fill_order = 1, sample_format = 1;
    bool alpha = false;
    std::vector<uint64_t> offsets, counts;
    std::vector<uint16_t> colormap;
};

struct Reader {
    const uint8_t* p;
    size_t n;
    bool be;
    bool ok(uint64_t off, uint64_t len) const { return off <= n && len <= n - off; }
    uint16_t u16(uint64_t o) const { return be ? (uint16_t)(p[o] << 8 | p[o + 1]) : (uint16_t)(p[o] | p[o + 1] << 8); }
    uint32_t u32(uint64_t o) const {
        return be ? ((uint32_t)p[o] << 24 | (uint32_t)p[o + 1] << 16 | (uint32_t)p[o + 2] << 8 | p[o + 3])
                  : ((uint32_t)p[o + 3] << 24 | (uint32_t)p[o + 2] << 16 | (uint32_t)p[o + 1] << 8 | p[o]);
    }
};

// Values of one IFD entry (types BYTE, SHORT, LONG) as 64-bit integers.
bool entry_values(const Reader& r, uint64_t e, std::vector<uint64_t>& out) {
    const uint16_t type = r.u16(e + 2);
    const uint32_t cnt = r.u32(e + 4);
    const uint32_t sz = (type == 1 || type == 6 || type == 7) ? 1 : (type == 3 || type == 8) ? 2 : (type == 4 || type == 9) ? 4 : 0;
    if (!sz) return false;
    const uint64_t bytes = (uint64_t)sz * cnt;
    uint64_t off = e + 8;
    if (bytes > 4) {
        off = r.u32(e + 8);
        if (!r.ok(off, bytes)) return false;
    }
    out.resize(cnt);
    for (uint32_t i = 0; i < cnt; i++) out[i] = sz == 1 ? r.p[off + i] : sz == 2 ? r.u16(off + 2 * i) : r.u32(off + 4ull * i);
    return true;
}

int parse_tiff(const uint8_t* buf, size_t n, TiffInfo& t) {
    if (n < 8) return ffail(UGSM_ERR_INVALID, "not a TIFF file (too short)");
    if (buf[0] == 'I' && buf[1] == 'I') t.be = false;
    else if (buf[0] == 'M' && buf[1] == 'M') t.be = true;
    else return ffail(UGSM_ERR_INVALID, "not a TIFF file (byte-order mark)");
    Reader r{buf, n, t.be};
    if (r.u16(2) != 42) return ffail(UGSM_ERR_INVALID, "not a classic TIFF (BigTIFF is not supported)");
    const uint64_t ifd = r.u32(4);
    if (!r.ok(ifd, 2)) return ffail(UGSM_ERR_INVALID, "TIFF: bad IFD offset");
    const uint16_t ne = r.u16(ifd);
    if (!r.ok(ifd + 2, 12ull * ne)) return ffail(UGSM_ERR_INVALID, "TIFF: truncated IFD");
    std::vector<uint64_t> v, bpsv, extra;
    bool tiled = false;
    for (uint16_t i = 0; i < ne; i++) {
        const uint64_t e = ifd + 2 + 12ull * i;
        const uint16_t tag = r.u16(e);
        switch (tag) {
            case 256: case 257: case 259: case 262: case 266: case 277: case 278: case 284: case 317: case 322:
            case 323: case 339:
                if (!entry_values(r, e, v) || v.empty()) return ffail(UGSM_ERR_INVALID, "TIFF: bad tag value");
                break;
            default: break;
        }
        switch (tag) {
            case 256: t.width = (uint32_t)v[0]; break;
            case 257: t.height = (uint32_t)v[0]; break;
            case 258: if (!entry_values(r, e, bpsv)) return ffail(UGSM_ERR_INVALID, "TIFF: bad BitsPerSample"); break;
            case 259: t.compression = (uint32_t)v[0]; break;
            case 262: t.photometric = (uint32_t)v[0]; break;
            case 266: t.fill_order = (uint32_t)v[0]; break;
            case 273: case 324: if (!entry_values(r, e, t.offsets)) return ffail(UGSM_ERR_INVALID, "TIFF: bad offsets"); tiled = tiled || tag == 324; break;
            case 277: t.spp = (uint32_t)v[0]; break;
            case 278: t.rows_per_strip = (uint32_t)v[0]; break;
            case 279: case 325: if (!entry_values(r, e, t.counts)) return ffail(UGSM_ERR_INVALID, "TIFF: bad byte counts"); break;
            case 284: t.planar = (uint32_t)v[0]; break;
            case 317: t.predictor = (uint32_t)v[0]; break;
            case 320: {
                std::vector<uint64_t> cm;
                if (!entry_values(r, e, cm)) return ffail(UGSM_ERR_INVALID, "TIFF: bad ColorMap");
                t.colormap.assign(cm.begin(), cm.end());
                break;
            }
            case 322: t.tile_w = (uint32_t)v[0]; break;
            case 323: t.tile_h = (uint32_t)v[0]; break;
            case 338: if (!entry_values(r, e, extra)) return ffail(UGSM_ERR_INVALID, "TIFF: bad ExtraSamples"); break;
            case 339: t.sample_format = (uint32_t)v[0]; break;
            default: break;
        }
    }
    if (!bpsv.empty()) {
        t.bps = (uint32_t)bpsv[0];
        for (uint64_t b : bpsv)
            if (b != t.bps) return ffail(UGSM_ERR_INVALID, "TIFF: mixed BitsPerSample are not supported");
    }
    for (uint64_t x : extra) t.alpha = t.alpha || x == 1 || x == 2;
    if (!tiled) t.tile_w = t.tile_h = 0;
    if (t.width == 0 || t.height == 0 || t.width > (1u << 20) || t.height > (1u << 20)) return ffail(UGSM_ERR_INVALID, "TIFF: bad image size");
    if (t.bps != 8 && t.bps != 16) return ffail(UGSM_ERR_INVALID, "TIFF: only 8- and 16-bit samples are supported");
    if (t.sample_format != 1) return ffail(UGSM_ERR_INVALID, "TIFF: only unsigned integer samples are supported");
    if (t.fill_order != 1) return ffail(UGSM_ERR_INVALID, "TIFF: FillOrder 2 is not supported");
    if (t.planar != 1 && t.planar != 2) return ffail(UGSM_ERR_INVALID, "TIFF: bad PlanarConfiguration");
    if (t.predictor != 1 && t.predictor != 2) return ffail(UGSM_ERR_INVALID, "TIFF: only the horizontal predictor is supported");
    if (t.compression != 1 && t.compression != 5 && t.compression != 32773 && t.compression != 8 && t.compression != 32946)
        return ffail(UGSM_ERR_INVALID, "TIFF: compression %s%ld is not supported (none, LZW, PackBits, Deflate are)", "", (long)t.compression);
    if (t.photometric > 3) return ffail(UGSM_ERR_INVALID, "TIFF: photometric interpretation %s%ld is not supported", "", (long)t.photometric);
    if (t.photometric == 2 && t.spp < 3) return ffail(UGSM_ERR_INVALID, "TIFF: RGB image with fewer than 3 samples");
    if (t.photometric == 3 && (t.bps != 8 || t.colormap.size() < 3 * 256)) return ffail(UGSM_ERR_INVALID, "TIFF: palette image needs 8 bits and a ColorMap");
    if (t.alpha) return ffail(UGSM_ERR_INVALID, "TIFF: images with an alpha channel are not supported");
    if (t.spp < 1 || t.spp > 8) return ffail(UGSM_ERR_INVALID, "TIFF: bad SamplesPerPixel");
    if (t.offsets.empty() || t.offsets.size() != t.counts.size()) return ffail(UGSM_ERR_INVALID, "TIFF: strip/tile tables missing");
    if (t.tile_w == 0 && t.rows_per_strip > t.height) t.rows_per_strip = t.height;
    if (t.tile_w == 0 && t.rows_per_strip == 0) return ffail(UGSM_ERR_INVALID, "TIFF: RowsPerStrip is zero");
    return UGSM_OK;
}

// TIFF LZW (MSB-first codes, 9..12 bits, "early change").  Returns the number of bytes written.  Strings are emitted
// back to back, so the string of a table entry -- the previous string plus the first byte of the current one -- is a
// contiguous run of the OUTPUT: an entry is (position, length) and decoding a code is a forward copy of earlier
// output in 8-byte pieces (the bytes a piece copies beyond the string's end are overwritten by the next string).  The
// 256 literals are a table in front of the output (K_LZW_PREFIX bytes, which the caller provides before `dst`, plus
// 16 bytes of slack behind `cap`), so literal and string codes take the same path: no data-dependent branch per code.
constexpr size_t K_LZW_PREFIX = 272;  // 256 literals + 8 bytes an 8-byte piece may read past the last one, 16-aligned
size_t lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    struct Ent { uint32_t pos, len; };  // string = base[pos .. pos + len), base = dst - K_LZW_PREFIX
    static thread_local Ent tab[4096 + 2];
    uint8_t* const base = dst - K_LZW_PREFIX;
    for (uint32_t i = 0; i < 256; i++) base[i] = (uint8_t)i, tab[i] = Ent{i, 1};
    uint32_t next = 258, bits = 9;
    uint64_t acc = 0;  // valid bits at the top
    int nacc = 0;
    size_t ip = 0;
    uint32_t op = (uint32_t)K_LZW_PREFIX;
    const uint32_t end = (uint32_t)(K_LZW_PREFIX + cap);
    uint32_t ppos = 0, plen = 0;  // the previous string
    bool have_prev = false;
    for (;;) {
        if (nacc <= 32) {
            if (ip + 4 <= n) {
                const uint32_t w = (uint32_t)src[ip] << 24 | (uint32_t)src[ip + 1] << 16 | (uint32_t)src[ip + 2] << 8 | src[ip + 3];
                acc |= (uint64_t)w << (32 - nacc);
                ip += 4, nacc += 32;
            } else {
                while (nacc <= 56 && ip < n) acc |= (uint64_t)src[ip++] << (56 - nacc), nacc += 8;
            }
        }
        if (nacc < (int)bits) break;
        const uint32_t code = (uint32_t)(acc >> (64 - bits));
        acc <<= bits;
        nacc -= (int)bits;
        if (code == 256) {
            next = 258, bits = 9, have_prev = false;
            continue;
        }
        if (code == 257 || op >= end) break;
        if (code > next || (code == next && !have_prev) || (code >= 258 && !have_prev)) break;  // corrupt stream
        uint32_t slen;
        if (code < 256) {  // literal: the common case on noisy images, kept off the table path
            base[op] = (uint8_t)code;
            slen = 1;
        } else {
            const bool kwkwk = (code == next);
            const uint32_t spos = kwkwk ? ppos : tab[code].pos;
            const uint32_t cl = kwkwk ? plen : tab[code].len;  // bytes copied from earlier output
            slen = cl + (kwkwk ? 1u : 0u);
            if (op + slen > end) {  // the last strip may carry padding: copy what fits
                for (uint32_t i = 0; op + i < end; i++) base[op + i] = (kwkwk && i == cl) ? base[ppos] : base[spos + i];
                op = end;
                break;
            }
            for (uint32_t i = 0; i < cl; i += 8) memcpy(base + op + i, base + spos + i, 8);
            if (kwkwk) base[op + cl] = base[ppos];
        }
        if (have_prev && next < 4096) {
            tab[next] = Ent{ppos, plen + 1};
            next++;
            if (next + 1 >= (1u << bits) && bits < 12) bits++;  // early change: switch one code early
        }
        ppos = op, plen = slen, have_prev = true;
        op += slen;
    }
    return op - K_LZW_PREFIX;
}

size_t packbits_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    size_t ip = 0, op = 0;
    while (ip < n && op < cap) {
        const int8_t c = (int8_t)src[ip++];
        if (c >= 0) {
            size_t k = (size_t)c + 1;
            if (k > n - ip) k = n - ip;
            if (k > cap - op) k = cap - op;
            memcpy(dst + op, src + ip, k);
            ip += k, op += k;
        } else if (c != -128) {
            if (ip >= n) break;
            size_t k = (size_t)(-c) + 1;
            if (k > cap - op) k = cap - op;
            memset(dst + op, src[ip++], k);
            op += k;
        }
    }
    return op;
}

size_t inflate_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    uLongf out = (uLongf)cap;
    const int rc = uncompress(dst, &out, src, (uLong)n);
    return (rc == Z_OK || rc == Z_BUF_ERROR) ? (size_t)out : 0;
}

// One strip / tile into `raw` (rows x row_bytes, samples as stored: big-endian 16-bit are swapped to host order).
int decode_chunk(const Reader& r, const TiffInfo& t, size_t idx, uint8_t* raw, size_t rows, size_t row_samples) {
    const size_t bytes = rows * row_samples * (t.bps / 8);
    const uint64_t off = t.offsets[idx], cnt = t.counts[idx];
    if (!r.ok(off, cnt)) return ffail(UGSM_ERR_INVALID, "TIFF: strip/tile %s%ld lies outside the file", "", (long)idx);
    size_t got;
    switch (t.compression) {
        case 1: got = cnt < bytes ? (size_t)cnt : bytes; memcpy(raw, r.p + off, got); break;
        case 5: got = lzw_decode(r.p + off, (size_t)cnt, raw, bytes); break;
        case 32773: got = packbits_decode(r.p + off, (size_t)cnt, raw, bytes); break;
        default: got = inflate_decode(r.p + off, (size_t)cnt, raw, bytes); break;
    }
    if (got < bytes) return ffail(UGSM_ERR_INVALID, "TIFF: strip/tile %s%ld is short", "", (long)idx);
    return UGSM_OK;
}

// `samples` consecutive samples per pixel in a decoded row: undo the horizontal predictor (in stored byte order for
// 8-bit, in value space for 16-bit) and bring 16-bit samples to host order.
void finish_row(const TiffInfo& t, uint8_t* row, size_t pixels, size_t samples) {
    if (t.bps == 8) {
        if (t.predictor == 2)
            for (size_t i = samples; i < pixels * samples; i++) row[i] = (uint8_t)(row[i] + row[i - samples]);
    } else {
        uint16_t* w = reinterpret_cast<uint16_t*>(row);
        const size_t n = pixels * samples;
        if (t.be)
            for (size_t i = 0; i < n; i++) w[i] = (uint16_t)(w[i] >> 8 | w[i] << 8);  // file order -> little-endian host
        if (t.predictor == 2)
            for (size_t i = samples; i < n; i++) w[i] = (uint16_t)(w[i] + w[i - samples]);
    }
}

inline uint8_t to8_rgb16(uint16_t v) { return (uint8_t)(((uint32_t)v * 255u + 32767u) / 65535u); }  // libtiff RGBA path
inline uint8_t to8_grey16(uint16_t v) { return (uint8_t)(v >> 8); }                                  // cv 16u -> 8u

// Pixels of one decoded row segment -> B,G,R bytes.  planes[s] points at sample s of the first pixel, `stride`
// samples apart (chunky: planes[s] = row + s, stride = spp; planar: separate rows, stride = 1).
void emit_bgr(const TiffInfo& t, const uint8_t* const planes[3], size_t stride, size_t pixels, uint8_t* dst) {
    if (t.bps == 8) {
        if (t.photometric == 2) {
            for (size_t i = 0; i < pixels; i++) {
                dst[3 * i + 0] = planes[2][i * stride];
                dst[3 * i + 1] = planes[1][i * stride];
                dst[3 * i + 2] = planes[0][i * stride];
            }
        } else if (t.photometric == 3) {
            for (size_t i = 0; i < pixels; i++) {
                const uint8_t k = planes[0][i * stride];
                dst[3 * i + 0] = (uint8_t)(t.colormap[512 + k] >> 8);
                dst[3 * i + 1] = (uint8_t)(t.colormap[256 + k] >> 8);
                dst[3 * i + 2] = (uint8_t)(t.colormap[k] >> 8);
            }
        } else {
            for (size_t i = 0; i < pixels; i++) {
                uint8_t g = planes[0][i * stride];
                if (t.photometric == 0) g = (uint8_t)(255 - g);
                dst[3 * i + 0] = dst[3 * i + 1] = dst[3 * i + 2] = g;
            }
        }
    } else {
        const uint16_t* const* p16 = reinterpret_cast<const uint16_t* const*>(planes);
        if (t.photometric == 2) {
            for (size_t i = 0; i < pixels; i++) {
                dst[3 * i + 0] = to8_rgb16(p16[2][i * stride]);
                dst[3 * i + 1] = to8_rgb16(p16[1][i * stride]);
                dst[3 * i + 2] = to8_rgb16(p16[0][i * stride]);
            }
        } else {
            for (size_t i = 0; i < pixels; i++) {
                uint16_t g = p16[0][i * stride];
                if (t.photometric == 0) g = (uint16_t)(65535 - g);
                dst[3 * i + 0] = dst[3 * i + 1] = dst[3 * i + 2] = to8_grey16(g);
            }
        }
    }
}

// R,G,B -> B,G,R for `pixels` pixels; 16 pixels (48 bytes) at a time through three 16-byte shuffles where SSSE3 is
// compiled in, 4 pixels per 12-byte word group otherwise.
void swap_rgb_row(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t pixels) {
    size_t i = 0;
#if defined(__SSSE3__)
    const __m128i m0a = _mm_setr_epi8(2, 1, 0, 5, 4, 3, 8, 7, 6, 11, 10, 9, 14, 13, 12, -1);
    const __m128i m0b = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 1);
    const __m128i m1a = _mm_setr_epi8(-1, 15, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m128i m1b = _mm_setr_epi8(0, -1, 4, 3, 2, 7, 6, 5, 10, 9, 8, 13, 12, 11, -1, 15);
    const __m128i m1c = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, -1);
    const __m128i m2a = _mm_setr_epi8(14, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m128i m2b = _mm_setr_epi8(-1, 3, 2, 1, 6, 5, 4, 9, 8, 7, 12, 11, 10, 15, 14, 13);
    for (; i + 16 <= pixels; i += 16) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 3 * i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 3 * i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 3 * i + 32));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + 3 * i), _mm_or_si128(_mm_shuffle_epi8(a, m0a), _mm_shuffle_epi8(b, m0b)));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + 3 * i + 16),
                         _mm_or_si128(_mm_or_si128(_mm_shuffle_epi8(a, m1a), _mm_shuffle_epi8(b, m1b)), _mm_shuffle_epi8(c, m1c)));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + 3 * i + 32), _mm_or_si128(_mm_shuffle_epi8(b, m2a), _mm_shuffle_epi8(c, m2b)));
    }
#endif
    for (; i < pixels; i++) {
        const uint8_t r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
        dst[3 * i] = b, dst[3 * i + 1] = g, dst[3 * i + 2] = r;
    }
}

int decode_tiff(const uint8_t* buf, size_t n, const TiffInfo& t, uint8_t* dst, size_t step) {
    Reader r{buf, n, t.be};
    // What cameras write: uncompressed 8-bit chunky RGB strips.  One pass from the mapped file into the destination.
    if (t.compression == 1 && t.bps == 8 && t.photometric == 2 && t.spp == 3 && t.planar == 1 && t.predictor == 1 &&
        t.tile_w == 0) {
        const size_t rps = t.rows_per_strip, nstrips = (t.height + rps - 1) / rps;
        if (t.offsets.size() < nstrips) return ffail(UGSM_ERR_INVALID, "TIFF: too few strips/tiles");
        for (size_t sidx = 0; sidx < nstrips; sidx++) {
            const size_t y0 = sidx * rps, rows = (y0 + rps <= t.height) ? rps : t.height - y0;
            const size_t bytes = rows * (size_t)t.width * 3;
            if (!r.ok(t.offsets[sidx], bytes) || t.counts[sidx] < bytes)
                return ffail(UGSM_ERR_INVALID, "TIFF: strip/tile %s%ld is short", "", (long)sidx);
            const uint8_t* src = buf + t.offsets[sidx];
            for (size_t y = 0; y < rows; y++) swap_rgb_row(src + y * (size_t)t.width * 3, dst + (y0 + y) * step, t.width);
        }
        return UGSM_OK;
    }
    const size_t bpsb = t.bps / 8;
    const bool tiled = t.tile_w != 0;
    const size_t cw = tiled ? t.tile_w : t.width;                   // chunk width in pixels
    const size_t ch = tiled ? t.tile_h : t.rows_per_strip;          // chunk height
    if (cw == 0 || ch == 0) return ffail(UGSM_ERR_INVALID, "TIFF: bad tile size");
    const size_t across = tiled ? (t.width + cw - 1) / cw : 1, down = (t.height + ch - 1) / ch;
    const size_t per_plane = across * down;
    const size_t nplanes = t.planar == 2 ? t.spp : 1;
    if (t.offsets.size() < per_plane * nplanes) return ffail(UGSM_ERR_INVALID, "TIFF: too few strips/tiles");
    const size_t csamples = t.planar == 2 ? 1 : t.spp;              // samples per pixel inside a chunk
    const size_t used = t.photometric == 2 ? 3 : 1;                 // samples that make the colour
    std::vector<uint8_t> raw[3];
    for (size_t cy = 0; cy < down; cy++) {
        const size_t y0 = cy * ch;
        const size_t rows_img = (y0 + ch <= t.height) ? ch : t.height - y0;  // rows of this chunk inside the image
        const size_t rows_dec = tiled ? ch : rows_img;                       // tiles are always stored whole
        for (size_t cx = 0; cx < across; cx++) {
            const size_t x0 = cx * cw, cols_img = (x0 + cw <= t.width) ? cw : t.width - x0;
            const size_t np = t.planar == 2 ? used : 1;
            for (size_t s = 0; s < np; s++) {
                raw[s].resize(K_LZW_PREFIX + rows_dec * cw * csamples * bpsb + 16);  // literal table in front, slack behind
                const int rc = decode_chunk(r, t, s * per_plane + cy * across + cx, raw[s].data() + K_LZW_PREFIX, rows_dec, cw * csamples);
                if (rc != UGSM_OK) return rc;
            }
            for (size_t y = 0; y < rows_img; y++) {
                const uint8_t* planes[3];
                const size_t row_off = y * cw * csamples * bpsb;
                if (t.planar == 2) {
                    for (size_t s = 0; s < used; s++) {
                        finish_row(t, raw[s].data() + K_LZW_PREFIX + row_off, cw, 1);
                        planes[s] = raw[s].data() + K_LZW_PREFIX + row_off;
                    }
                } else {
                    finish_row(t, raw[0].data() + K_LZW_PREFIX + row_off, cw, t.spp);
                    for (size_t s = 0; s < used; s++) planes[s] = raw[0].data() + K_LZW_PREFIX + row_off + s * bpsb;
                }
                emit_bgr(t, planes, csamples, cols_img, dst + (y0 + y) * step + 3 * x0);
            }
        }
    }
    return UGSM_OK;
}

// A file mapped read-only (no copy of the 48 MB a 16 MP image occupies; pages come straight from the page cache).
struct MappedFile {
    const uint8_t* p = nullptr;
    size_t n = 0;
    ~MappedFile() {
        if (p && n) munmap(const_cast<uint8_t*>(p), n);
    }
    int open(const char* path) {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return ffail(UGSM_ERR_INVALID, "cannot open %s", path);
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size <= 0) {
            ::close(fd);
            return ffail(UGSM_ERR_INVALID, "cannot size %s", path);
        }
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (m == MAP_FAILED) return ffail(UGSM_ERR_INVALID, "cannot map %s", path);
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        p = static_cast<const uint8_t*>(m);
        n = (size_t)st.st_size;
        return UGSM_OK;
    }
};

int read_file(const char* path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return ffail(UGSM_ERR_INVALID, "cannot open %s", path);
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) {
        fclose(f);
        return ffail(UGSM_ERR_INVALID, "cannot size %s", path);
    }
    out.resize((size_t)sz);
    const size_t got = sz ? fread(out.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) return ffail(UGSM_ERR_INVALID, "short read on %s", path);
    return UGSM_OK;
}

// ------------------------------------------------------------------------------------ image list (XML)
// The subset of OpenCV's FileStorage XML that readStringList consumes: the first element inside <opencv_storage> is
// a whitespace-separated sequence of strings; entries may be quoted; comments are skipped; &amp; &lt; &gt; &quot;
// &apos; are unescaped.
int read_list(const char* path, std::vector<std::string>& out) {
    std::vector<uint8_t> raw;
    int rc = read_file(path, raw);
    if (rc != UGSM_OK) return rc;
    std::string s(raw.begin(), raw.end());
    for (size_t a; (a = s.find("<!--")) != std::string::npos;) {
        const size_t b = s.find("-->", a + 4);
        s.erase(a, b == std::string::npos ? std::string::npos : b + 3 - a);
    }
    size_t p = s.find("<opencv_storage");
    if (p == std::string::npos) return ffail(UGSM_ERR_INVALID, "%s is not an OpenCV storage file", path);
    p = s.find('>', p);
    if (p == std::string::npos) return ffail(UGSM_ERR_INVALID, "%s: malformed XML", path);
    const size_t open = s.find('<', p + 1);
    if (open == std::string::npos || s[open + 1] == '/') return ffail(UGSM_ERR_INVALID, "%s: no top-level node", path);
    const size_t open_end = s.find('>', open);
    if (open_end == std::string::npos) return ffail(UGSM_ERR_INVALID, "%s: malformed XML", path);
    std::string name = s.substr(open + 1, open_end - open - 1);
    const size_t sp = name.find_first_of(" \t\r\n/");
    if (sp != std::string::npos) name.erase(sp);
    const size_t close = s.find("</" + name, open_end);
    if (close == std::string::npos) return ffail(UGSM_ERR_INVALID, "%s: unterminated sequence node", path);
    std::string body = s.substr(open_end + 1, close - open_end - 1);
    if (body.find('<') != std::string::npos) return ffail(UGSM_ERR_INVALID, "%s: first top-level node is not a sequence", path);
    out.clear();
    size_t i = 0;
    auto unescape = [](std::string v) {
        static const char* const ent[5][2] = {{"&amp;", "&"}, {"&lt;", "<"}, {"&gt;", ">"}, {"&quot;", "\""}, {"&apos;", "'"}};
        for (auto& e : ent)
            for (size_t q; (q = v.find(e[0])) != std::string::npos;) v.replace(q, strlen(e[0]), e[1]);
        return v;
    };
    while (i < body.size()) {
        while (i < body.size() && isspace((unsigned char)body[i])) i++;
        if (i >= body.size()) break;
        std::string tok;
        if (body[i] == '"') {
            const size_t e = body.find('"', i + 1);
            tok = body.substr(i + 1, e == std::string::npos ? std::string::npos : e - i - 1);
            i = e == std::string::npos ? body.size() : e + 1;
        } else {
            size_t e = i;
            while (e < body.size() && !isspace((unsigned char)body[e])) e++;
            tok = body.substr(i, e - i);
            i = e;
        }
        out.push_back(unescape(tok));
    }
    return UGSM_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------- the ring
struct ugsm_feeder {
    std::vector<std::string> list;
    std::string base_dir;
    long pairs = 0;  // pairs to deliver in total
    int rows = 0, cols = 0;
    size_t step = 0;
    bool pinned = false;
    struct Slot {
        uint8_t *L = nullptr, *R = nullptr;
        long index = -1;     // pair held (or being decoded)
        int state = 0;       // 0 free, 1 decoding, 2 ready, 3 handed out
        int rc = UGSM_OK;
        std::string err;
    };
    std::vector<Slot> slots;
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv;
    long next_decode = 0, next_out = 0;
    bool closing = false;
    double decode_ms = 0;  // summed over images
    char last_error[512] = "";

    std::string path_of(long image) const {  // Settings::nextImage (settings.h:30-45): past the end -> next-to-last entry
        long k = image;
        const long nlist = (long)list.size();
        if (k >= nlist) k = nlist - 2 + ((image - nlist) & 1);
        const std::string& p = list[(size_t)k];
        if (!p.empty() && p[0] != '/' && !base_dir.empty()) return base_dir + "/" + p;
        return p;
    }
};

namespace {

int load_image(const std::string& path, uint8_t* dst, size_t step, int rows, int cols) {
    MappedFile file;
    int rc = file.open(path.c_str());
    if (rc != UGSM_OK) return rc;
    TiffInfo t;
    rc = parse_tiff(file.p, file.n, t);
    if (rc != UGSM_OK) return rc;
    if ((int)t.height != rows || (int)t.width != cols) return ffail(UGSM_ERR_INVALID, "%s does not have the size of the first image of the list", path.c_str());
    return decode_tiff(file.p, file.n, t, dst, step);
}

void decode_loop(ugsm_feeder* f) {
    for (;;) {
        long i;
        ugsm_feeder::Slot* s;
        {
            std::unique_lock<std::mutex> lk(f->mu);
            // pairs are claimed in order, a pair's slot is index % slots: wait until its previous tenant is released
            f->cv.wait(lk, [&] {
                return f->closing || f->next_decode >= f->pairs ||
                       f->slots[(size_t)(f->next_decode % (long)f->slots.size())].state == 0;
            });
            if (f->closing || f->next_decode >= f->pairs) return;
            i = f->next_decode++;
            s = &f->slots[(size_t)(i % (long)f->slots.size())];
            s->state = 1, s->index = i;
        }
        struct timespec a, b;
        clock_gettime(CLOCK_MONOTONIC, &a);
        int rc;
        try {  // nothing may escape a decode thread
            rc = load_image(f->path_of(2 * i), s->L, f->step, f->rows, f->cols);
            if (rc == UGSM_OK) rc = load_image(f->path_of(2 * i + 1), s->R, f->step, f->rows, f->cols);
        } catch (...) {
            rc = ffail(UGSM_ERR_NOMEM, "out of memory while decoding pair %s%ld", "", i);
        }
        clock_gettime(CLOCK_MONOTONIC, &b);
        {
            std::lock_guard<std::mutex> lk(f->mu);
            s->rc = rc;
            s->err = rc == UGSM_OK ? "" : g_feeder_error;
            s->state = 2;
            f->decode_ms += (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6;
        }
        f->cv.notify_all();
    }
}

}  // namespace

extern "C" {

const char* ugsm_feeder_last_error(const ugsm_feeder* f) { return f ? f->last_error : g_feeder_error; }

int ugsm_tiff_info(const char* path, int* rows, int* cols, int* samples, int* bits) {
    if (!path) return UGSM_ERR_INVALID;
    try {
        MappedFile file;
        int rc = file.open(path);
        if (rc != UGSM_OK) return rc;
        TiffInfo t;
        rc = parse_tiff(file.p, file.n, t);
        if (rc != UGSM_OK) return rc;
        if (rows) *rows = (int)t.height;
        if (cols) *cols = (int)t.width;
        if (samples) *samples = (int)t.spp;
        if (bits) *bits = (int)t.bps;
        return UGSM_OK;
    } catch (...) {
        return ffail(UGSM_ERR_NOMEM, "ugsm_tiff_info: out of memory");
    }
}

int ugsm_tiff_read_bgr8(const char* path, uint8_t* dst, size_t step, int rows, int cols) {
    if (!path || !dst || rows < 1 || cols < 1 || step < (size_t)cols * 3) return UGSM_ERR_INVALID;
    try {
        return load_image(path, dst, step, rows, cols);
    } catch (...) {
        return ffail(UGSM_ERR_NOMEM, "ugsm_tiff_read_bgr8: out of memory");
    }
}

int ugsm_feeder_open(const char* list_xml, long pairs, int slots, int threads, ugsm_feeder** out) {
    if (!list_xml || !out || slots < 1 || threads < 1 || slots > 1024 || threads > 256) return UGSM_ERR_INVALID;
    *out = nullptr;
    ugsm_feeder* f = nullptr;
    try {
        f = new ugsm_feeder;
        int rc = read_list(list_xml, f->list);
        if (rc == UGSM_OK && (f->list.size() < 2)) rc = ffail(UGSM_ERR_INVALID, "%s lists fewer than two images", list_xml);
        if (rc != UGSM_OK) {
            delete f;
            return rc;
        }
        const std::string lp(list_xml);
        const size_t sl = lp.rfind('/');
        f->base_dir = sl == std::string::npos ? "." : lp.substr(0, sl);
        f->pairs = pairs > 0 ? pairs : (long)(f->list.size() / 2);
        int samples, bits;
        rc = ugsm_tiff_info(f->path_of(0).c_str(), &f->rows, &f->cols, &samples, &bits);
        if (rc != UGSM_OK) {
            delete f;
            return rc;
        }
        f->step = (size_t)f->cols * 3;
        const size_t bytes = f->step * f->rows;
        f->slots.resize((size_t)slots);
        // page-locked slots; without a CUDA device the ring still works with ordinary memory (ugsm_match_* stages
        // pageable buffers itself)
        f->pinned = true;
        for (auto& s : f->slots) {
            void *l = nullptr, *r = nullptr;
            if (ugsm_host_alloc(bytes, &l) != UGSM_OK || ugsm_host_alloc(bytes, &r) != UGSM_OK) {
                if (l) ugsm_host_free(l);
                f->pinned = false;
                break;
            }
            s.L = (uint8_t*)l, s.R = (uint8_t*)r;
        }
        if (!f->pinned) {
            for (auto& s : f->slots) {
                if (s.L) ugsm_host_free(s.L);
                if (s.R) ugsm_host_free(s.R);
                s.L = s.R = nullptr;
            }
            const size_t padded = (bytes + 4095) / 4096 * 4096;
            for (auto& s : f->slots) {
                s.L = (uint8_t*)aligned_alloc(4096, padded);
                s.R = (uint8_t*)aligned_alloc(4096, padded);
                if (!s.L || !s.R) {
                    ugsm_feeder_close(f);
                    return ffail(UGSM_ERR_NOMEM, "ugsm_feeder_open: out of host memory");
                }
            }
        }
        for (int i = 0; i < threads; i++) f->threads.emplace_back(decode_loop, f);
        *out = f;
        return UGSM_OK;
    } catch (...) {
        if (f) ugsm_feeder_close(f);
        return ffail(UGSM_ERR_NOMEM, "ugsm_feeder_open: out of memory");
    }
}

int ugsm_feeder_info(const ugsm_feeder* f, ugsm_feeder_info_t* out) {
    if (!f || !out) return UGSM_ERR_INVALID;
    out->rows = f->rows, out->cols = f->cols, out->step = f->step;
    out->list_images = (long)f->list.size();
    out->pairs = f->pairs;
    out->slots = (int)f->slots.size();
    out->pinned = f->pinned ? 1 : 0;
    {
        std::lock_guard<std::mutex> lk(const_cast<ugsm_feeder*>(f)->mu);
        out->decode_ms = f->decode_ms;
    }
    return UGSM_OK;
}

int ugsm_feeder_next(ugsm_feeder* f, ugsm_feeder_pair* out) {
    if (!f || !out) return UGSM_ERR_INVALID;
    std::unique_lock<std::mutex> lk(f->mu);
    if (f->next_out >= f->pairs) return UGSM_FEEDER_END;
    const long i = f->next_out++;
    ugsm_feeder::Slot& s = f->slots[(size_t)(i % (long)f->slots.size())];
    f->cv.wait(lk, [&] { return f->closing || (s.index == i && s.state == 2); });
    if (f->closing) return UGSM_ERR_INVALID;
    s.state = 3;
    out->left = s.L, out->right = s.R;
    out->rows = f->rows, out->cols = f->cols, out->step = f->step;
    out->index = i;
    out->slot = (int)(i % (long)f->slots.size());
    if (s.rc != UGSM_OK) {
        snprintf(f->last_error, sizeof f->last_error, "pair %ld: %s", i, s.err.c_str());
        snprintf(g_feeder_error, sizeof g_feeder_error, "%s", f->last_error);
        const int rc = s.rc;
        s.state = 0, s.index = -1;  // a failed pair holds no data: its slot is free again
        lk.unlock();
        f->cv.notify_all();
        return rc;
    }
    return UGSM_OK;
}

int ugsm_feeder_release(ugsm_feeder* f, int slot) {
    if (!f || slot < 0 || slot >= (int)f->slots.size()) return UGSM_ERR_INVALID;
    {
        std::lock_guard<std::mutex> lk(f->mu);
        if (f->slots[(size_t)slot].state != 3) return UGSM_ERR_INVALID;
        f->slots[(size_t)slot].state = 0;
        f->slots[(size_t)slot].index = -1;
    }
    f->cv.notify_all();
    return UGSM_OK;
}

int ugsm_feeder_close(ugsm_feeder* f) {
    if (!f) return UGSM_OK;
    {
        std::lock_guard<std::mutex> lk(f->mu);
        f->closing = true;
    }
    f->cv.notify_all();
    for (auto& t : f->threads) t.join();
    for (auto& s : f->slots) {
        if (f->pinned) {
            ugsm_host_free(s.L);
            ugsm_host_free(s.R);
        } else {
            free(s.L);
            free(s.R);
        }
    }
    delete f;
    return UGSM_OK;
}

}  // extern "C"
