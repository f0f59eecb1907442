// write_image (src/renderprocess.rs:1501-1530): sRGB gamma (misc.rs:46-52), `clamp(255 g + 0.5, 0, 255) as u8`,
// RGBA with alpha 255, saved as PNG.  The reference delegates the container to the `image` crate
// (0.23.14); here a minimal encoder writes the same pixels with stored (uncompressed) deflate
// blocks — any PNG reader decodes identical bytes.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace rrt {

inline double gamma_correct(double v) {  // misc.rs:46-52
    return v <= 0.0031308 ? 12.92 * v : 1.055 * std::pow(v, 1.0 / 2.4) - 0.055;
}
inline uint8_t to_u8(double v) {  // `clamp_t(255.0 * gamma_correct(v) + 0.5, 0.0, 255.0) as u8` (NaN -> 0)
    double x = 255.0 * gamma_correct(v) + 0.5;
    if (!(x == x)) return 0;
    x = x < 0.0 ? 0.0 : (x > 255.0 ? 255.0 : x);
    return (uint8_t)x;
}
inline void rgb_to_rgba8(const double* rgb, size_t npix, uint8_t* out) {
    for (size_t i = 0; i < npix; ++i) {
        out[4 * i] = to_u8(rgb[3 * i]);
        out[4 * i + 1] = to_u8(rgb[3 * i + 1]);
        out[4 * i + 2] = to_u8(rgb[3 * i + 2]);
        out[4 * i + 3] = 255;
    }
}

inline uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

inline bool write_png_rgba8(const std::string& path, const uint8_t* rgba, uint32_t w, uint32_t h) {
    std::vector<uint8_t> raw;  // filter byte 0 + scanline
    raw.reserve((size_t)h * (4 * (size_t)w + 1));
    for (uint32_t y = 0; y < h; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), rgba + 4 * (size_t)y * w, rgba + 4 * (size_t)(y + 1) * w);
    }
    std::vector<uint8_t> z = {0x78, 0x01};  // zlib header, no compression
    uint32_t a = 1, b = 0;                  // Adler-32
    size_t pos = 0;
    while (pos < raw.size() || raw.empty()) {
        size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xFF));
        z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xFF));
        z.push_back((uint8_t)((~n >> 8) & 0xFF));
        for (size_t i = 0; i < n; ++i) {
            a = (a + raw[pos + i]) % 65521u;
            b = (b + a) % 65521u;
        }
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
        if (raw.empty()) break;
    }
    const uint32_t adler = (b << 16) | a;
    for (int k = 3; k >= 0; --k) z.push_back((uint8_t)(adler >> (8 * k)));
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    auto be32 = [](uint32_t v, uint8_t* o) {
        o[0] = (uint8_t)(v >> 24);
        o[1] = (uint8_t)(v >> 16);
        o[2] = (uint8_t)(v >> 8);
        o[3] = (uint8_t)v;
    };
    auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
        uint8_t len[4], crcb[4];
        be32((uint32_t)data.size(), len);
        std::fwrite(len, 1, 4, f);
        std::fwrite(type, 1, 4, f);
        if (!data.empty()) std::fwrite(data.data(), 1, data.size(), f);
        uint32_t crc = crc32_update(0, reinterpret_cast<const uint8_t*>(type), 4);
        if (!data.empty()) crc = crc32_update(crc, data.data(), data.size());
        be32(crc, crcb);
        std::fwrite(crcb, 1, 4, f);
    };
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::fwrite(sig, 1, 8, f);
    std::vector<uint8_t> ihdr(13);
    be32(w, &ihdr[0]);
    be32(h, &ihdr[4]);
    ihdr[8] = 8;   // bit depth
    ihdr[9] = 6;   // RGBA
    ihdr[10] = ihdr[11] = ihdr[12] = 0;
    chunk("IHDR", ihdr);
    chunk("IDAT", z);
    chunk("IEND", {});
    const bool ok = std::ferror(f) == 0;
    std::fclose(f);
    return ok;
}

}  // namespace rrt
