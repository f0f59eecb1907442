// See image_host.hpp.
#include "image_host.hpp"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace rrt {

namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

bool read_png_rgb8(const std::string& path, Image8* out, std::string* err) {
    auto fail = [&](const std::string& m) {
        if (err) *err = path + ": " + m;
        return false;
    };
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail("cannot open");
    std::vector<uint8_t> data;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
    std::fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (data.size() < 8 || std::memcmp(data.data(), sig, 8) != 0) return fail("not a PNG file (only PNG images are read)");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette;
    size_t pos = 8;
    bool end = false;
    while (!end && pos + 12 <= data.size()) {
        const uint32_t len = be32(&data[pos]);
        const char* type = reinterpret_cast<const char*>(&data[pos + 4]);
        if (pos + 12 + (size_t)len > data.size()) return fail("truncated chunk");
        const uint8_t* body = &data[pos + 8];
        if (std::memcmp(type, "IHDR", 4) == 0 && len >= 13) {
            w = be32(body);
            h = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
        } else if (std::memcmp(type, "PLTE", 4) == 0) {
            palette.assign(body, body + len);
        } else if (std::memcmp(type, "IDAT", 4) == 0) {
            idat.insert(idat.end(), body, body + len);
        } else if (std::memcmp(type, "IEND", 4) == 0) {
            end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (w == 0 || h == 0 || w > 16384 || h > 16384) return fail("image size out of range");
    if (interlace != 0) return fail("interlaced PNGs are not read");
    int channels = 0;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: return fail("unknown colour type");
    }
    if (ctype == 3 ? !(depth == 1 || depth == 2 || depth == 4 || depth == 8) : depth != 8)
        return fail("only 8 bits per channel (palette: 1 to 8) are read");
    const size_t bpp_bits = (size_t)channels * (size_t)depth;
    const size_t stride = ((size_t)w * bpp_bits + 7) / 8, bpp = std::max<size_t>(1, bpp_bits / 8);
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf raw_len = (uLongf)raw.size();
    const int zr = uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || raw_len != raw.size()) return fail("inflate failed or the image data has the wrong size");
    // undo the scanline filters in place
    std::vector<uint8_t> prev(stride, 0);
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t* row = &raw[(stride + 1) * (size_t)y];
        const int ft = row[0];
        uint8_t* cur = row + 1;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = cur[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: return fail("unknown scanline filter");
            }
            cur[i] = (uint8_t)v;
        }
        std::memcpy(prev.data(), cur, stride);
    }
    out->width = w;
    out->height = h;
    out->rgb.assign((size_t)w * h * 3, 0);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* cur = &raw[(stride + 1) * (size_t)y + 1];
        uint8_t* o = &out->rgb[(size_t)y * w * 3];
        for (uint32_t x = 0; x < w; ++x) {
            uint8_t r, g, b;
            if (ctype == 2 || ctype == 6) {
                r = cur[(size_t)x * channels];
                g = cur[(size_t)x * channels + 1];
                b = cur[(size_t)x * channels + 2];
            } else if (ctype == 0 || ctype == 4) {
                r = g = b = cur[(size_t)x * channels];
            } else {
                const size_t bit = (size_t)x * depth;
                const uint32_t idx = (cur[bit / 8] >> (8 - depth - (bit % 8))) & ((1u << depth) - 1u);
                if ((size_t)idx * 3 + 2 >= palette.size()) return fail("palette index out of range");
                r = palette[idx * 3];
                g = palette[idx * 3 + 1];
                b = palette[idx * 3 + 2];
            }
            o[3 * x] = r;
            o[3 * x + 1] = g;
            o[3 * x + 2] = b;
        }
    }
    return true;
}

std::vector<double> mip_weight_lut() {
    std::vector<double> lut(128);
    for (int i = 0; i < 128; ++i) {
        const double alpha = 2.0, r2 = (double)i / 127.0;
        lut[i] = std::exp(-alpha * r2) - std::exp(-alpha);
    }
    return lut;
}

MipView HostMipMap::host_view(const double* weight_lut) const {
    MipView v{};
    v.n_levels = (uint32_t)levels.size();
    for (size_t i = 0; i < levels.size(); ++i) {
        v.level[i].data = levels[i].data();
        v.level[i].u_res = u_res[i];
        v.level[i].v_res = v_res[i];
        v.level[i].u_blocks = ((u_res[i] + 3) & ~(uint64_t)3) >> 2;
        v.level[i].cells = levels[i].size() / 3;
    }
    v.wrap = wrap;
    v.trilinear = trilinear;
    v.max_aniso = max_aniso;
    v.weight_lut = weight_lut;
    return v;
}

namespace {
struct Px {
    double c[3];
};
double lanczos(double x, double tau) {  // texture/mod.rs:191-204
    x = std::fabs(x);
    if (x < 1e-5) return 1.0;
    if (x > 1.0) return 0.0;
    x *= kPi;
    const double s = std::sin(x * tau) / (x * tau);
    const double l = std::sin(x) / x;
    return s * l;
}
struct Weights {
    uint64_t first;
    double w[4];
};
std::vector<Weights> resample_weights(uint64_t old_res, uint64_t new_res) {  // mipmap.rs:24-46
    std::vector<Weights> out(new_res);
    for (uint64_t i = 0; i < new_res; ++i) {
        const double center = ((double)i + 0.5) * (double)old_res / (double)new_res;
        Weights r;
        r.first = f64_as_usize(std::floor(center - 2.0 + 0.5));  // Q33: a negative first texel becomes 0
        for (int j = 0; j < 4; ++j) r.w[j] = lanczos(((double)(r.first + (uint64_t)j) + 0.5 - center) / 2.0, 2.0);
        const double inv = 1.0 / (r.w[0] + r.w[1] + r.w[2] + r.w[3]);
        for (double& x : r.w) x *= inv;
        out[i] = r;
    }
    return out;
}
uint64_t wrap_index(uint64_t i, uint64_t n, uint32_t wrap) {
    if (wrap == MIPWRAP_REPEAT) return i - (i / n) * n;
    if (wrap == MIPWRAP_CLAMP) return i > n - 1 ? n - 1 : i;
    return i;
}
uint64_t round_pow2(uint64_t v) {  // misc.rs:318-330
    v -= 1;
    v |= v >> 1;
    v |= v >> 2;
    v |= v >> 4;
    v |= v >> 8;
    v |= v >> 16;
    return v + 1;
}
}  // namespace

bool make_mipmap(const Image8& img, bool trilinear, double max_aniso, uint32_t wrap, HostMipMap* out, std::string* err) {
    const uint64_t rx = img.width, ry = img.height;
    if (rx == 0 || ry == 0 || wrap > MIPWRAP_CLAMP) {
        if (err) *err = "empty image or unknown wrap mode";
        return false;
    }
    // texels: value / 255, rows flipped (renderprocess.rs:543-559)
    std::vector<Px> tex(rx * ry);
    for (uint64_t y = 0; y < ry; ++y)
        for (uint64_t x = 0; x < rx; ++x) {
            const uint8_t* p = &img.rgb[3 * (y * rx + x)];
            tex[(ry - 1 - y) * rx + x] = Px{{(double)p[0] / 255.0, (double)p[1] / 255.0, (double)p[2] / 255.0}};
        }
    uint64_t px = rx, py = ry;
    if ((rx & (rx - 1)) != 0 || (ry & (ry - 1)) != 0) {
        px = round_pow2(rx);
        py = round_pow2(ry);
        std::vector<Px> zoom(px * py, Px{{0, 0, 0}});
        const auto sw = resample_weights(rx, px);
        for (uint64_t t = 0; t < ry; ++t)
            for (uint64_t s = 0; s < px; ++s) {
                Px acc{{0, 0, 0}};
                for (uint64_t j = 0; j < 4; ++j) {
                    const uint64_t o = wrap_index(sw[s].first + j, rx, wrap);
                    if (o < rx)
                        for (int k = 0; k < 3; ++k) acc.c[k] += tex[t * rx + o].c[k] * sw[s].w[j];
                }
                zoom[t * px + s] = acc;
            }
        const auto tw = resample_weights(ry, py);
        std::vector<Px> work(py);
        for (uint64_t s = 0; s < px; ++s) {
            for (uint64_t t = 0; t < py; ++t) {
                Px acc{{0, 0, 0}};
                for (uint64_t j = 0; j < 4; ++j) {
                    const uint64_t o = wrap_index(tw[t].first + j, ry, wrap);
                    if (o < ry)
                        for (int k = 0; k < 3; ++k) acc.c[k] += zoom[o * px + s].c[k] * tw[t].w[j];
                }
                work[t] = acc;
            }
            for (uint64_t t = 0; t < py; ++t)
                for (int k = 0; k < 3; ++k) {  // Spectrum::clamp(0, inf) via clamp_t: a NaN passes through
                    const double v = work[t].c[k];
                    zoom[t * px + s].c[k] = v < 0.0 ? 0.0 : v;
                }
        }
        tex.swap(zoom);
    }
    out->levels.clear();
    out->u_res.clear();
    out->v_res.clear();
    out->wrap = wrap;
    out->trilinear = trilinear ? 1 : 0;
    out->max_aniso = max_aniso;
    auto new_level = [&](uint64_t ur, uint64_t vr) -> bool {
        const uint64_t ru = (ur + 3) & ~(uint64_t)3, rv = (vr + 3) & ~(uint64_t)3;
        // every offset the index formula produces must stay inside the storage (Q31), or the reference panics
        const uint64_t ub = ru >> 2;
        uint64_t max_u = 0, max_v = 0;
        for (uint64_t u = ur > 4 ? ur - 4 : 0; u < ur; ++u) max_u = std::max(max_u, 16 * (u & 3) + (u >> 2));
        for (uint64_t v = vr > 4 ? vr - 4 : 0; v < vr; ++v) max_v = std::max(max_v, 16 * ub * (v & 3) + 4 * (v >> 2));
        if (max_u + max_v >= ru * rv) {
            if (err) *err = "an image this small indexes outside its BlockedArray (the reference panics, memory.rs:76-85)";
            return false;
        }
        out->levels.emplace_back(ru * rv * 3, 0.0);
        out->u_res.push_back(ur);
        out->v_res.push_back(vr);
        return true;
    };
    if (!new_level(px, py)) return false;
    {
        std::vector<double>& d = out->levels[0];
        const uint64_t ub = ((px + 3) & ~(uint64_t)3) >> 2;
        for (uint64_t u = 0; u < px; ++u)      // BlockedArray::new: u outer, v inner — the LAST write to a cell stays
            for (uint64_t v = 0; v < py; ++v) {
                const uint64_t o = blocked_offset(ub, u, v);
                for (int k = 0; k < 3; ++k) d[3 * o + k] = tex[v * px + u].c[k];
            }
    }
    const std::vector<double> lut = mip_weight_lut();
    const uint64_t n_levels = 1 + f64_as_usize(std::log2((double)std::max(px, py)));
    for (uint64_t i = 1; i < n_levels; ++i) {
        const uint64_t sr = std::max<uint64_t>(out->u_res[i - 1] / 2, 1), tr = std::max<uint64_t>(out->v_res[i - 1] / 2, 1);
        if (std::min(sr, tr) < 64) break;
        if (out->levels.size() >= (size_t)kMipMaxLevels) {
            if (err) *err = "image too large: more than " + std::to_string(kMipMaxLevels) + " MIPMap levels";
            return false;
        }
        if (!new_level(sr, tr)) return false;
        const MipView view = out->host_view(lut.data());  // the finer level is read through MIPMap::texel
        std::vector<double>& d = out->levels[i];
        const uint64_t ub = ((sr + 3) & ~(uint64_t)3) >> 2;
        for (uint64_t t = 0; t < tr; ++t)
            for (uint64_t s = 0; s < sr; ++s) {
                const Rgb c = (mip_texel(view, (uint32_t)(i - 1), 2 * s, 2 * t) + mip_texel(view, (uint32_t)(i - 1), 2 * s + 1, 2 * t) +
                               mip_texel(view, (uint32_t)(i - 1), 2 * s, 2 * t + 1) + mip_texel(view, (uint32_t)(i - 1), 2 * s + 1, 2 * t + 1)) * 0.25;
                const uint64_t o = blocked_offset(ub, s, t);
                d[3 * o] = c.r;
                d[3 * o + 1] = c.g;
                d[3 * o + 2] = c.b;
            }
    }
    return true;
}

Dist2DView HostDist2D::host_view() const {
    Dist2DView v{};
    v.func = func.data();
    v.cdf = cdf.data();
    v.func_int = func_int.data();
    v.mcdf = mcdf.data();
    v.m_func_int = m_func_int;
    v.nu = nu;
    v.nv = nv;
    return v;
}

namespace {
// Distribution1D::new (sampling.rs:17-40)
double make_cdf(const double* f, uint32_t n, double* cdf) {
    cdf[0] = 0.0;
    for (uint32_t i = 1; i <= n; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / (double)n;
    const double func_int = cdf[n];
    if (func_int == 0.0) {
        for (uint32_t i = 1; i <= n; ++i) cdf[i] = (double)i / (double)n;
    } else {
        for (uint32_t i = 1; i <= n; ++i) cdf[i] /= func_int;
    }
    return func_int;
}
}  // namespace

void make_env_distribution(const MipView& lmap, HostDist2D* out) {
    const uint64_t width = 2 * lmap.level[0].u_res, height = 2 * lmap.level[0].v_res;
    out->nu = (uint32_t)width;
    out->nv = (uint32_t)height;
    out->func.resize(width * height);
    const double fwidth = 0.5 / std::fmin((double)width, (double)height);
    for (uint64_t v = 0; v < height; ++v) {
        const double vp = ((double)v + 0.5) / (double)height;
        const double sin_theta = std::sin(kPi * ((double)v + 0.5) / (double)height);
        for (uint64_t u = 0; u < width; ++u) {
            const double up = ((double)u + 0.5) / (double)width;
            out->func[v * width + u] = lum(mip_lookup_w(lmap, P2{up, vp}, fwidth)) * sin_theta;
        }
    }
    out->cdf.resize((width + 1) * height);
    out->func_int.resize(height);
    for (uint64_t v = 0; v < height; ++v)
        out->func_int[v] = make_cdf(&out->func[v * width], (uint32_t)width, &out->cdf[v * (width + 1)]);
    out->mcdf.resize(height + 1);
    out->m_func_int = make_cdf(out->func_int.data(), (uint32_t)height, out->mcdf.data());
}

}  // namespace rrt
