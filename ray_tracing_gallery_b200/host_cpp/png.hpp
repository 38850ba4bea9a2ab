// png.hpp — PNG -> RGBA8 with the semantics of the reference's image path:
// `image::load_from_memory_with_format(bytes, Png)?.to_rgba8()` (src/util_functions.rs:247-249).
// Non-interlaced PNGs of every colour type / bit depth; 16-bit samples are narrowed by `>> 8`
// (the rule of `image` 0.23; the blue-noise asset is 16-bit grey).  zlib does the inflate.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace b200rt_host {

struct ImageRgba8 {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> texels;  // width * height * 4
};

namespace png_detail {
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace png_detail

inline ImageRgba8 decode_png_rgba8(const uint8_t* data, size_t size) {
    using namespace png_detail;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || std::memcmp(data, sig, 8) != 0) throw std::runtime_error("PNG: bad signature");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t p = 8;
    bool end = false;
    while (!end && p + 12 <= size) {
        uint32_t len = be32(data + p);
        const uint8_t* type = data + p + 4;
        const uint8_t* body = data + p + 8;
        if (p + 12 + (size_t)len > size) throw std::runtime_error("PNG: truncated chunk");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) throw std::runtime_error("PNG: bad IHDR");
            w = be32(body); h = be32(body + 4);
            depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) end = true;
        p += 12 + (size_t)len;
    }
    if (!w || !h) throw std::runtime_error("PNG: no IHDR");
    if (interlace) throw std::runtime_error("PNG: interlaced images are not supported");
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels) throw std::runtime_error("PNG: bad colour type");
    // legal bit depths per colour type (PNG spec 11.2.2): anything else would shift by garbage or divide by zero below
    const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                     : (depth == 8 || depth == 16);
    if (!depth_ok) throw std::runtime_error("PNG: bit depth not allowed for this colour type");
    size_t bpp_bits = (size_t)channels * depth;
    size_t stride = (w * bpp_bits + 7) / 8;
    size_t bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;  // filter unit in bytes
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf out_len = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || out_len != raw.size()) throw std::runtime_error("PNG: inflate failed");
    // unfilter in place
    std::vector<uint8_t> prev(stride, 0);
    std::vector<uint8_t> pix(stride * (size_t)h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* src = raw.data() + (stride + 1) * (size_t)y;
        uint8_t* cur = pix.data() + stride * (size_t)y;
        int ft = src[0];
        src++;
        for (size_t i = 0; i < stride; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = src[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: throw std::runtime_error("PNG: bad filter type");
            }
            cur[i] = (uint8_t)v;
        }
        std::memcpy(prev.data(), cur, stride);
    }
    // expand to RGBA8
    ImageRgba8 img;
    img.width = w; img.height = h;
    img.texels.resize((size_t)w * h * 4);
    auto sample = [&](const uint8_t* row, size_t index) -> uint32_t {  // `index`-th sample of the row at `depth` bits
        if (depth == 8) return row[index];
        if (depth == 16) return ((uint32_t)row[2 * index] << 8) | row[2 * index + 1];
        size_t bit = index * depth;
        return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
    };
    auto to8 = [&](uint32_t v) -> uint8_t {
        if (depth == 8) return (uint8_t)v;
        if (depth == 16) return (uint8_t)(v >> 8);
        return (uint8_t)(v * 255u / ((1u << depth) - 1u));  // 1/2/4-bit greys scale to 0..255
    };
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* row = pix.data() + stride * (size_t)y;
        uint8_t* out = img.texels.data() + (size_t)y * w * 4;
        for (uint32_t x = 0; x < w; x++, out += 4) {
            switch (ctype) {
                case 0: {
                    uint32_t g = sample(row, x);
                    out[0] = out[1] = out[2] = to8(g);
                    out[3] = (trns.size() >= 2 && g == (((uint32_t)trns[0] << 8) | trns[1])) ? 0 : 255;
                    break;
                }
                case 2: {
                    uint32_t r = sample(row, 3 * (size_t)x), g = sample(row, 3 * (size_t)x + 1), b = sample(row, 3 * (size_t)x + 2);
                    out[0] = to8(r); out[1] = to8(g); out[2] = to8(b);
                    // tRNS colour key (three 16-bit samples): that exact colour is fully transparent, like image's to_rgba8
                    const bool key = trns.size() >= 6 && r == (((uint32_t)trns[0] << 8) | trns[1]) && g == (((uint32_t)trns[2] << 8) | trns[3]) &&
                                     b == (((uint32_t)trns[4] << 8) | trns[5]);
                    out[3] = key ? 0 : 255;
                    break;
                }
                case 3: {
                    uint32_t i = sample(row, x);
                    if (3 * (size_t)i + 2 >= plte.size()) throw std::runtime_error("PNG: palette index out of range");
                    out[0] = plte[3 * i]; out[1] = plte[3 * i + 1]; out[2] = plte[3 * i + 2];
                    out[3] = i < trns.size() ? trns[i] : 255;
                    break;
                }
                case 4:
                    out[0] = out[1] = out[2] = to8(sample(row, 2 * (size_t)x));
                    out[3] = to8(sample(row, 2 * (size_t)x + 1));
                    break;
                default:
                    for (int k = 0; k < 4; k++) out[k] = to8(sample(row, 4 * (size_t)x + k));
            }
        }
    }
    return img;
}

}  // namespace b200rt_host
