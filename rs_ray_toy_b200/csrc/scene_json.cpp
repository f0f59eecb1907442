// scene.json + Wavefront .obj loading: the subset of src/renderprocess.rs (deploy_render,
// make_scene, make_materials, make_triangle_mesh, make_all_lights, make_aggregate, make_film,
// make_camera, make_sampler, make_integrator) and src/objparser.rs that the hot path consumes
// (SURVEY.md Appendix C).  Keys, defaults and quirks follow the reference:
//   * objs[].world_pos/rotation/scale are parsed but never applied to vertices (Q7);
//   * a material parameter is a texture NAME; the only textures honoured here are BilerpTextures
//     whose corners agree (the loader reads v10 and v11 from key "v01", renderprocess.rs:326-329),
//     i.e. constants; everything else falls back to the documented default or is refused;
//   * point lights ignore their transform (Q17); `Filter` must exist (may be {}).
#include "scene_json.hpp"

#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "json_min.hpp"

namespace rrt {
namespace {

using json::Value;

std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
// read_i64 / read_f64 / read_bool / read_string (renderprocess.rs:136-169)
double read_f64(const Value& v, const char* key, double def) {
    const Value* x = v.get(key);
    return (x && x->is_number()) ? x->num : def;
}
int64_t read_i64(const Value& v, const char* key, int64_t def) {
    const Value* x = v.get(key);
    return (x && x->is_number()) ? (int64_t)x->num : def;
}
bool read_bool(const Value& v, const char* key, bool def) {
    const Value* x = v.get(key);
    return (x && x->kind == Value::Bool) ? x->b : def;
}
std::string read_string(const Value& v, const char* key, const char* def) {
    const Value* x = v.get(key);
    return (x && x->is_string()) ? x->str : std::string(def);
}
bool read_xyz(const Value& v, const char* key, double out[3], double dx, double dy, double dz) {
    out[0] = dx;
    out[1] = dy;
    out[2] = dz;
    const Value* x = v.get(key);
    if (x && x->is_array() && x->arr.size() >= 3 && x->arr[0]->is_number() && x->arr[1]->is_number() && x->arr[2]->is_number()) {
        for (int k = 0; k < 3; ++k) out[k] = x->arr[k]->num;
        return true;
    }
    return false;
}
// make_spectrum (renderprocess.rs:1055-1076): { "values": [r, g, b] }
void read_spectrum(const Value& v, const char* key, double out[3], double def) {
    out[0] = out[1] = out[2] = def;
    const Value* x = v.get(key);
    if (!x) return;
    const Value* vals = x->get("values");
    if (!vals) return;
    if (!vals->is_array() || vals->arr.size() < 3) throw std::runtime_error(std::string("Failed to parse Spectrum ") + key);
    for (int k = 0; k < 3; ++k) out[k] = vals->arr[k]->num;
}

void mul44(const double a[4][4], const double b[4][4], double r[4][4]) {  // transform.rs:165-177
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] + a[i][3] * b[3][j];
}
void normalize3(double v[3]) {  // Vector3f::normalize (geometry.rs:925-931)
    double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (l == 0.0) return;
    for (int k = 0; k < 3; ++k) v[k] /= l;
}
// make_to_world (renderprocess.rs:242-252): translate * rotate(angle, normalize(axis)) * scale
Transform make_to_world(const Value& cfg) {
    double pos[3], axis[3], sc[3];
    read_xyz(cfg, "world_pos", pos, 0, 0, 0);
    read_xyz(cfg, "rotation_axis", axis, 0, 0, 0);
    read_xyz(cfg, "scale", sc, 1, 1, 1);
    normalize3(axis);
    const double angle = read_f64(cfg, "rotation_angle", 0.0);
    double T[4][4] = {{1, 0, 0, pos[0]}, {0, 1, 0, pos[1]}, {0, 0, 1, pos[2]}, {0, 0, 0, 1}};
    double Ti[4][4] = {{1, 0, 0, -pos[0]}, {0, 1, 0, -pos[1]}, {0, 0, 1, -pos[2]}, {0, 0, 0, 1}};
    // Transform::rotate (transform.rs:327-351) normalises the axis again
    double a[3] = {axis[0], axis[1], axis[2]};
    normalize3(a);
    const double rad = (3.14159265358979323846264338327950288 / 180.0) * angle;
    const double s = std::sin(rad), c = std::cos(rad);
    double R[4][4] = {{a[0] * a[0] + (1.0 - a[0] * a[0]) * c, a[0] * a[1] * (1.0 - c) - a[2] * s, a[0] * a[2] * (1.0 - c) + a[1] * s, 0},
                      {a[0] * a[1] * (1.0 - c) + a[2] * s, a[1] * a[1] + (1.0 - a[1] * a[1]) * c, a[1] * a[2] * (1.0 - c) - a[0] * s, 0},
                      {a[0] * a[2] * (1.0 - c) - a[1] * s, a[1] * a[2] * (1.0 - c) + a[0] * s, a[2] * a[2] + (1.0 - a[2] * a[2]) * c, 0},
                      {0, 0, 0, 1}};
    double Ri[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) Ri[i][j] = R[j][i];
    double S[4][4] = {{sc[0], 0, 0, 0}, {0, sc[1], 0, 0}, {0, 0, sc[2], 0}, {0, 0, 0, 1}};
    double Si[4][4] = {{1.0 / sc[0], 0, 0, 0}, {0, 1.0 / sc[1], 0, 0}, {0, 0, 1.0 / sc[2], 0}, {0, 0, 0, 1}};
    Transform out;
    double TR[4][4], RiTi[4][4];
    mul44(T, R, TR);
    mul44(TR, S, out.m.m);
    // (T*R)*S inverse = S^-1 * (R^-1 * T^-1)  (Transform::mul, transform.rs:441-449)
    mul44(Ri, Ti, RiTi);
    mul44(Si, RiTi, out.inv.m);
    return out;
}

// objparser.rs:83-247
void parse_obj(const std::string& path, TriangleMesh* mesh) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<uint32_t> vi, ni, uvi;
    std::string line;
    auto parse_index = [](const std::string& tok, int64_t out[3]) {
        out[0] = out[1] = out[2] = -1;
        size_t start = 0;
        for (int k = 0; k < 3 && start <= tok.size(); ++k) {
            size_t slash = tok.find('/', start);
            std::string part = tok.substr(start, slash == std::string::npos ? std::string::npos : slash - start);
            bool digits = !part.empty();
            for (char c : part) digits &= (c >= '0' && c <= '9');
            if (digits) {
                int64_t v = std::strtoll(part.c_str(), nullptr, 10);
                out[k] = v > 0 ? v - 1 : -1;  // usize::from_str(idx) - 1 (:219)
            }
            if (slash == std::string::npos) break;
            start = slash + 1;
        }
    };
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::string tag;
        if (!(ss >> tag)) continue;
        if (tag == "v") {
            double x, y, z;
            if (!(ss >> x >> y >> z)) throw std::runtime_error("ParseObjError: bad vertex in " + path);
            mesh->p.insert(mesh->p.end(), {x, y, z});
        } else if (tag == "vt") {
            double u, v = 0.0;
            if (!(ss >> u)) throw std::runtime_error("ParseObjError: bad uv in " + path);
            ss >> v;
            mesh->uv.insert(mesh->uv.end(), {u, v});
        } else if (tag == "vn") {
            double x, y, z;
            if (!(ss >> x >> y >> z)) throw std::runtime_error("ParseObjError: bad normal in " + path);
            const double l = std::sqrt(x * x + y * y + z * z);  // Normal3f::normalize on read (:130)
            mesh->n.insert(mesh->n.end(), {x / l, y / l, z / l});
        } else if (tag == "f") {
            std::string t[3];
            if (!(ss >> t[0] >> t[1] >> t[2])) throw std::runtime_error("ParseObjError: Failed to get face element");
            int64_t e[3][3];
            for (int k = 0; k < 3; ++k) parse_index(t[k], e[k]);
            if (e[0][0] >= 0 && e[1][0] >= 0 && e[2][0] >= 0) {
                for (int k = 0; k < 3; ++k) vi.push_back((uint32_t)e[k][0]);
                const int64_t nuv = (int64_t)mesh->uv.size() / 2, nn = (int64_t)mesh->n.size() / 3;
                if (e[0][1] >= 0 && e[1][1] >= 0 && e[2][1] >= 0 && nuv > 0 && e[0][1] < nuv && e[1][1] < nuv && e[2][1] < nuv)
                    for (int k = 0; k < 3; ++k) uvi.push_back((uint32_t)e[k][1]);
                if (e[0][2] >= 0 && e[1][2] >= 0 && e[2][2] >= 0 && nn > 0 && e[0][2] < nn && e[1][2] < nn && e[2][2] < nn)
                    for (int k = 0; k < 3; ++k) ni.push_back((uint32_t)e[k][2]);
            }
        }
        // "#" and anything else: ignored (the reference logs unknown elements)
    }
    mesh->vi = vi;
    // index arrays that do not cover every face cannot be addressed per triangle (triangle.rs:88-98)
    if (ni.size() == vi.size()) mesh->ni = ni;
    if (uvi.size() == vi.size()) mesh->uvi = uvi;
    for (uint32_t x : mesh->vi)
        if ((size_t)x >= mesh->p.size() / 3) throw std::runtime_error("ParseObjError: vertex index out of range in " + path);
}

// make_textures (renderprocess.rs:298-515) flattened into the rrt_texture table: float textures first, then rgb
// textures, each in definition order.  Names are looked up in the map being filled, so a texture only sees the
// ones defined before it; an unknown child name becomes a constant (get_text_fallback, :282-296); a redefined name
// replaces the earlier entry for later lookups (HashMap::insert).
struct TextureTable {
    std::vector<rrt_texture> rows;
    std::vector<std::string> image_files;  // as written in the scene file; rrt_texture::t1 of an image row indexes this
    std::map<std::string, int32_t> f, rgb;  // name -> row; -2 = defined, but of a type outside the hot-path scope

    int32_t push(uint32_t kind) {
        rrt_texture t;
        std::memset(&t, 0, sizeof(t));
        t.kind = kind;
        t.mapping = RRT_TEXMAP_UV;
        t.t1 = t.t2 = t.amount = -1;
        t.map[0] = t.map[1] = 1.0;  // UVMapping2D::new(1, 1, 0, 0) when no mapping block is given (:607-609)
        for (int k = 0; k < 4; ++k) t.world_to_texture[5 * k] = 1.0;
        if (rows.size() >= RRT_MAX_TEXTURES) throw std::runtime_error("more than RRT_MAX_TEXTURES textures (constants included)");
        rows.push_back(t);
        return (int32_t)rows.size() - 1;
    }
    int32_t constant(const double v[3]) {
        const int32_t i = push(RRT_TEX_CONSTANT);
        for (int k = 0; k < 3; ++k) rows[i].v[0][k] = v[k];
        return i;
    }
    int32_t child(const std::map<std::string, int32_t>& names, const std::string& name, double def, bool is_rgb) {
        auto it = names.find(name);
        if (it != names.end()) {
            if (it->second == -2) throw std::runtime_error("texture '" + name + "' has a type outside the hot-path scope");
            return it->second;
        }
        const double v[3] = {def, is_rgb ? def : 0.0, is_rgb ? def : 0.0};
        return constant(v);
    }
    // make_texture_mapping_2d (:571-612)
    void mapping(int32_t i, const Value& tc, const Transform& to_world) {
        const Value* mc = tc.get("mapping");
        if (!mc) return;
        rrt_texture& t = rows[i];
        const std::string kind = read_string(*mc, "mapping", "uv");
        if (kind == "uv") {  // du / dv default to 1 once a mapping block is given (:582-583)
            t.mapping = RRT_TEXMAP_UV;
            t.map[0] = read_f64(*mc, "su", 1.0);
            t.map[1] = read_f64(*mc, "sv", 1.0);
            t.map[2] = read_f64(*mc, "du", 1.0);
            t.map[3] = read_f64(*mc, "dv", 1.0);
        } else if (kind == "planar") {
            t.mapping = RRT_TEXMAP_PLANAR;
            read_xyz(*mc, "v1", t.map, 1, 0, 0);
            read_xyz(*mc, "v2", t.map + 3, 0, 1, 0);
            t.map[6] = read_f64(*mc, "udelta", 0.0);
            t.map[7] = read_f64(*mc, "vdelta", 0.0);
        } else if (kind == "spherical" || kind == "cylindrical") {  // Transform::inverse(to_world)
            t.mapping = kind == "spherical" ? RRT_TEXMAP_SPHERICAL : RRT_TEXMAP_CYLINDRICAL;
            std::memcpy(t.world_to_texture, to_world.inv.m, sizeof(t.world_to_texture));
        } else {
            throw std::runtime_error("Unsupported Mapping Type " + kind);  // the reference panics (:601-606)
        }
    }
    void add(const Value& tc, bool is_rgb) {
        std::map<std::string, int32_t>& names = is_rgb ? rgb : f;
        const Transform to_world = make_to_world(tc);
        const std::string type = read_string(tc, "texture_type", "");
        const std::string name = read_string(tc, "texture_name", "DefaultTextureName");
        auto value = [&](const char* key, double def, double out[3]) {
            if (is_rgb) {
                read_spectrum(tc, key, out, def);
            } else {
                out[0] = read_f64(tc, key, def);
                out[1] = out[2] = 0.0;
            }
        };
        const std::string t1 = read_string(tc, "t1", "ErrorTextureName"), t2 = read_string(tc, "t2", "ErrorTextureName");
        int32_t i;
        if (type == "BilerpTexture") {
            double v[4][3];  // v10 and v11 are read from the key "v01" (:326-329, :439-442)
            value("v00", 0.0, v[0]);
            value("v01", 1.0, v[1]);
            value("v01", 0.0, v[2]);
            value("v01", 1.0, v[3]);
            bool same = true;
            for (int k = 1; k < 4; ++k)
                for (int c = 0; c < 3; ++c) same &= v[k][c] == v[0][c];
            if (same) {  // a constant in all but the 1-ulp sum of the four weights
                names[name] = constant(v[0]);
                return;
            }
            i = push(RRT_TEX_BILERP);
            std::memcpy(rows[i].v, v, sizeof(v));
            mapping(i, tc, to_world);
        } else if (type == "UVTexture" && is_rgb) {
            i = push(RRT_TEX_UV);
            mapping(i, tc, to_world);
        } else if (type == "ScaleTexture") {
            const int32_t c1 = child(names, t1, 1.0, is_rgb), c2 = child(names, t2, 1.0, is_rgb);
            i = push(RRT_TEX_SCALE);
            rows[i].t1 = c1;
            rows[i].t2 = c2;
        } else if (type == "MixTexture") {  // the amount is looked up under the key "t2" too (:319, :411)
            const int32_t c1 = child(names, t1, 0.0, is_rgb), c2 = child(names, t2, 1.0, is_rgb), amt = child(f, t2, 0.5, false);
            i = push(RRT_TEX_MIX);
            rows[i].t1 = c1;
            rows[i].t2 = c2;
            rows[i].amount = amt;
        } else if (type == "CheckerBoardTexture") {
            const int64_t dim = read_i64(tc, "dimension", 2);
            if (dim != 2 && dim != 3) return;  // logged and skipped (:341-344)
            const int32_t c1 = child(names, t1, 1.0, is_rgb), c2 = child(names, t2, 0.0, is_rgb);
            if (dim == 2) {
                i = push(RRT_TEX_CHECKER2D);
                rows[i].aa = read_string(tc, "aamode", "closedform") == "none" ? 0u : 1u;  // AAMethod (:357-366)
                mapping(i, tc, to_world);
            } else {  // IdentityMapping3D::new(to_world): the matrix is used as world_to_texture as it stands
                i = push(RRT_TEX_CHECKER3D);
                std::memcpy(rows[i].world_to_texture, to_world.m.m, sizeof(rows[i].world_to_texture));
            }
            rows[i].t1 = c1;
            rows[i].t2 = c2;
        } else if (type == "WindyTexture" || type == "WrinkledTexture") {  // IdentityMapping3D::new(to_world) (:376-388, :485-497)
            i = push(type == "WindyTexture" ? RRT_TEX_WINDY : RRT_TEX_WRINKLED);
            std::memcpy(rows[i].world_to_texture, to_world.m.m, sizeof(rows[i].world_to_texture));
            if (type == "WrinkledTexture") {
                rows[i].map[0] = (double)(uint64_t)read_i64(tc, "octaves", 8);
                rows[i].map[1] = read_f64(tc, "omega", 0.5);
            }
        } else if (type == "ImageTexture" && is_rgb) {  // make_tex_info + load_image (:517-566); `gamma` and `scale` are read
            i = push(RRT_TEX_IMAGE);                    // and never used there
            rows[i].aa = read_bool(tc, "do_trilinear", false) ? 1u : 0u;
            rows[i].v[0][0] = read_f64(tc, "max_aniso", 8.0);
            const std::string wrap = read_string(tc, "wrap", "repeat");
            rows[i].v[0][1] = wrap == "black" ? (double)RRT_WRAP_BLACK : (wrap == "clamp" ? (double)RRT_WRAP_CLAMP : (double)RRT_WRAP_REPEAT);
            rows[i].t1 = (int32_t)image_files.size();
            image_files.push_back(read_string(tc, "filename", "DefaultTexture"));
            mapping(i, tc, to_world);
        } else if (type == "ImageTexture" || (type == "UVTexture" && !is_rgb)) {
            return;  // not a float texture type: "Unsupported Texture Type", nothing inserted
        } else {
            return;  // "Unsupported Texture Type": nothing inserted
        }
        names[name] = i;
    }
    // -> texture index, or -1 with the constant written to `out`
    int32_t resolve(int32_t i, const std::string& name, double out[3]) const {
        if (i == -2) throw std::runtime_error("texture '" + name + "' has a type outside the hot-path scope");
        if (rows[i].kind != RRT_TEX_CONSTANT) return i;
        for (int k = 0; k < 3; ++k) out[k] = rows[i].v[0][k];
        return -1;
    }
};
TextureTable collect_textures(const Value& root) {
    TextureTable t;
    if (const Value* a = root.get("float_texture"); a && a->is_array())
        for (const auto& tc : a->arr) t.add(*tc, false);
    if (const Value* a = root.get("rgb_texture"); a && a->is_array())
        for (const auto& tc : a->arr) t.add(*tc, true);
    return t;
}
// fetch_float_texture (:614-626): a name that is not a float texture panics in the reference
double tex_f(const Value& m, const TextureTable& t, const char* key, double def, int32_t* slot) {
    *slot = -1;
    const Value* x = m.get(key);
    if (x && x->is_string()) {
        auto it = t.f.find(x->str);
        if (it == t.f.end()) throw std::runtime_error("float texture '" + x->str + "' does not exist");
        double v[3] = {def, 0.0, 0.0};
        *slot = t.resolve(it->second, x->str, v);
        return v[0];
    }
    return def;
}
// fetch_rgb_texture (:643-661) falls through to the default when the name is unknown
void tex_rgb(const Value& m, const TextureTable& t, const char* key, double out[3], const double def[3], int32_t* slot) {
    *slot = -1;
    for (int k = 0; k < 3; ++k) out[k] = def[k];
    const Value* x = m.get(key);
    if (x && x->is_string()) {
        auto it = t.rgb.find(x->str);
        if (it != t.rgb.end()) *slot = t.resolve(it->second, x->str, out);
    }
}

// RGB of the reference's default copper spectra: material/metal.rs COPPER_N / COPPER_K pushed
// through RGBSpectrum::from_sampled (spectrum.rs:2701-2727); derivation in tests/golden/make_copper_rgb.py
const double kCopperN[3] = {0.19998972096819712, 0.922085788777433, 1.0998762520488314};
const double kCopperK[3] = {3.9046381767086675, 2.4476332238684626, 2.1376510366555137};

bool make_material(const Value& m, const TextureTable& t, rrt_material* out, int32_t slots[RRT_MATERIAL_SLOTS]) {
    const std::string type = read_string(m, "material_type", "");
    rrt_material r;
    std::memset(&r, 0, sizeof(r));
    for (int k = 0; k < RRT_MATERIAL_SLOTS; ++k) slots[k] = -1;
    r.u_roughness = r.v_roughness = -1.0;
    r.remap_roughness = read_bool(m, "remap_roughness", false) ? 1 : 0;
    const double c5[3] = {0.5, 0.5, 0.5}, c25[3] = {0.25, 0.25, 0.25}, c9[3] = {0.9, 0.9, 0.9}, c1[3] = {1.0, 1.0, 1.0};
    if (type == "MatteMaterial") {
        r.kind = RRT_MAT_MATTE;
        tex_rgb(m, t, "kd", r.kd, c5, &slots[RRT_SLOT_KD]);
        r.sigma = tex_f(m, t, "sigma", 0.0, &slots[RRT_SLOT_SIGMA]);
    } else if (type == "PlasticMaterial") {
        r.kind = RRT_MAT_PLASTIC;
        tex_rgb(m, t, "kd", r.kd, c25, &slots[RRT_SLOT_KD]);
        tex_rgb(m, t, "ks", r.ks, c25, &slots[RRT_SLOT_KS]);
        r.roughness = tex_f(m, t, "roughness", 0.1, &slots[RRT_SLOT_ROUGHNESS]);
    } else if (type == "MetalMaterial") {
        r.kind = RRT_MAT_METAL;
        tex_rgb(m, t, "eta", r.metal_eta, kCopperN, &slots[RRT_SLOT_METAL_ETA]);
        tex_rgb(m, t, "k", r.metal_k, kCopperK, &slots[RRT_SLOT_METAL_K]);
        r.roughness = tex_f(m, t, "roughness", 0.01, &slots[RRT_SLOT_ROUGHNESS]);
        r.u_roughness = tex_f(m, t, "u_roughness", -1.0, &slots[RRT_SLOT_U_ROUGHNESS]);
        r.v_roughness = tex_f(m, t, "v_roughness", -1.0, &slots[RRT_SLOT_V_ROUGHNESS]);
    } else if (type == "MirrorMaterial") {
        r.kind = RRT_MAT_MIRROR;
        tex_rgb(m, t, "kr", r.kr, c9, &slots[RRT_SLOT_KR]);
    } else if (type == "GlassMaterial") {
        r.kind = RRT_MAT_GLASS;
        tex_rgb(m, t, "kr", r.kr, c1, &slots[RRT_SLOT_KR]);
        tex_rgb(m, t, "kt", r.kt, c1, &slots[RRT_SLOT_KT]);
        r.eta = tex_f(m, t, "eta", 1.5, &slots[RRT_SLOT_ETA]);
        r.u_roughness = tex_f(m, t, "u_roughness", 0.0, &slots[RRT_SLOT_U_ROUGHNESS]);
        r.v_roughness = tex_f(m, t, "v_roughness", 0.0, &slots[RRT_SLOT_V_ROUGHNESS]);
    } else if (type == "TranslucentMaterial") {  // renderprocess.rs:695-720
        r.kind = RRT_MAT_TRANSLUCENT;
        tex_rgb(m, t, "kd", r.kd, c25, &slots[RRT_SLOT_KD]);
        tex_rgb(m, t, "ks", r.ks, c25, &slots[RRT_SLOT_KS]);
        r.roughness = tex_f(m, t, "roughness", 0.1, &slots[RRT_SLOT_ROUGHNESS]);
        tex_rgb(m, t, "reflect", r.kr, c25, &slots[RRT_SLOT_KR]);
        tex_rgb(m, t, "transmit", r.kt, c25, &slots[RRT_SLOT_KT]);
    } else if (type == "DisneyMaterial") {  // renderprocess.rs:810-860
        r.kind = RRT_MAT_DISNEY;
        const double c0[3] = {0.0, 0.0, 0.0};
        tex_rgb(m, t, "color", r.kd, c5, &slots[RRT_SLOT_KD]);
        r.metallic = tex_f(m, t, "metallic", 0.0, &slots[RRT_SLOT_METALLIC]);
        r.eta = tex_f(m, t, "eta", 1.5, &slots[RRT_SLOT_ETA]);
        r.roughness = tex_f(m, t, "roughness", 0.5, &slots[RRT_SLOT_ROUGHNESS]);
        r.specular_tint = tex_f(m, t, "specular_tint", 0.0, &slots[RRT_SLOT_SPECULAR_TINT]);
        r.anisotropic = tex_f(m, t, "anisotropic", 0.0, &slots[RRT_SLOT_ANISOTROPIC]);
        r.sheen = tex_f(m, t, "sheen", 0.0, &slots[RRT_SLOT_SHEEN]);
        r.sheen_tint = tex_f(m, t, "sheen_tint", 0.5, &slots[RRT_SLOT_SHEEN_TINT]);
        r.clearcoat = tex_f(m, t, "clearcoat", 0.0, &slots[RRT_SLOT_CLEARCOAT]);
        r.clearcoat_gloss = tex_f(m, t, "clearcoat_gloss", 1.0, &slots[RRT_SLOT_CLEARCOAT_GLOSS]);
        r.spec_trans = tex_f(m, t, "spec_trans", 0.0, &slots[RRT_SLOT_SPEC_TRANS]);
        tex_rgb(m, t, "scatter_distance", r.scatter_distance, c0, &slots[RRT_SLOT_SCATTER_DISTANCE]);
        r.thin = read_bool(m, "thin", false) ? 1 : 0;
        r.flatness = tex_f(m, t, "flatness", 0.0, &slots[RRT_SLOT_FLATNESS]);
        r.diff_trans = tex_f(m, t, "diff_trans", 1.0, &slots[RRT_SLOT_DIFF_TRANS]);
    } else if (type == "Debug") {  // renderprocess.rs:861-863: no parameters, no bump map
        r.kind = RRT_MAT_DEBUG;
        r.remap_roughness = 0;
        *out = r;
        return true;
    } else {
        return false;  // "Unsupported Material Type" (renderprocess.rs:864-866): no entry; MixMaterial is handled by the caller
    }
    if (const Value* bm = m.get("bump_map"); bm && bm->is_string()) {  // fetch_float_texture_opt(.., "bump_map", None)
        auto it = t.f.find(bm->str);
        if (it == t.f.end()) throw std::runtime_error("float texture '" + bm->str + "' does not exist");
        if (it->second == -2) throw std::runtime_error("texture '" + bm->str + "' has a type outside the hot-path scope");
        slots[RRT_SLOT_BUMP_MAP] = it->second;  // the table row itself, constant or not
    }
    *out = r;
    return true;
}

}  // namespace

void load_scene_json(const std::string& path, const std::string& overrides_json, uint64_t seed, LoadedScene* out) {
    json::ValuePtr root = json::parse(read_file(path));
    if (!root->is_object()) throw std::runtime_error("scene.json: top level must be an object");
    if (!overrides_json.empty()) {
        json::ValuePtr ov = json::parse(overrides_json);
        if (!ov->is_object()) throw std::runtime_error("overrides must be a JSON object");
        for (const auto& kv : ov->obj) root->set(kv.first, kv.second);
    }
    std::string dir = ".";
    if (size_t slash = path.find_last_of('/'); slash != std::string::npos) dir = path.substr(0, slash);

    const TextureTable tex = collect_textures(*root);
    out->textures = tex.rows;
    // ---- make_materials ----
    std::map<std::string, uint32_t> material_index;
    if (const Value* a = root->get("materials"); a && a->is_array())
        for (const auto& mc : a->arr) {
            rrt_material m;
            int32_t slots[RRT_MATERIAL_SLOTS];
            if (read_string(*mc, "material_type", "") == "MixMaterial" && material_index.count(read_string(*mc, "mat1", "")) &&
                material_index.count(read_string(*mc, "mat2", "")))
                // renderprocess.rs:681-692 indexes scene_global.materials, still empty while make_materials runs (Q25)
                throw std::runtime_error("MixMaterial over two existing materials: the reference panics while loading (Q25)");
            if (make_material(*mc, tex, &m, slots)) {
                material_index[read_string(*mc, "material_name", "DefaultMaterialName")] = (uint32_t)out->materials.size();
                out->materials.push_back(m);
                out->material_slots.insert(out->material_slots.end(), slots, slots + RRT_MATERIAL_SLOTS);
            }
        }
    // ---- make_triangle_mesh ----
    std::map<std::string, uint32_t> mesh_index;
    if (const Value* a = root->get("objs"); a && a->is_array())
        for (const auto& oc : a->arr) {
            TriangleMesh mesh;
            const std::string file = read_string(*oc, "filename", "DefaultObj");
            parse_obj(file.size() && file[0] == '/' ? file : dir + "/" + file, &mesh);
            mesh_index[read_string(*oc, "obj_name", "DefaultObjName")] = (uint32_t)out->scene.meshes.size();
            out->scene.meshes.push_back(std::move(mesh));
        }
    // ---- make_aggregate ----
    const Value* agg = root->get("Aggregate");
    if (!agg) throw std::runtime_error("No Aggregate Config Defined");
    out->max_prims_in_node = (uint32_t)read_i64(*agg, "max_prims_in_node", 4);
    if (const Value* prims = agg->get("primitives"); prims && prims->is_array())
        for (const auto& pc : prims->arr) {
            const std::string type = read_string(*pc, "primitive_type", "");
            auto mit = material_index.find(read_string(*pc, "material_name", "DefaultMaterialName"));
            const Value* inst = pc->get("instances");
            const bool instanced = inst && inst->is_array();
            if (type == "sphere") {
                if (mit == material_index.end()) continue;
                Sphere s;
                s.obj_to_world = make_to_world(*pc);
                s.radius = read_f64(*pc, "radius", 1.0);
                s.z_min = read_f64(*pc, "z_min", -s.radius);
                s.z_max = read_f64(*pc, "z_max", s.radius);
                s.phi_max_deg = read_f64(*pc, "phi_max", 360.0);
                out->scene.spheres.push_back(s);
                const uint32_t sid = (uint32_t)out->scene.spheres.size() - 1;
                if (instanced) {
                    for (const auto& ic : inst->arr) {
                        out->scene.instances.push_back(make_to_world(*ic));
                        out->scene.prims.push_back({SHAPE_SPHERE, sid, 0, (int32_t)out->scene.instances.size() - 1, mit->second});
                    }
                } else {
                    out->scene.prims.push_back({SHAPE_SPHERE, sid, 0, -1, mit->second});
                }
            } else if (type == "triangle") {
                auto oit = mesh_index.find(read_string(*pc, "obj_name", "DefaultObjName"));
                if (oit == mesh_index.end() || mit == material_index.end()) continue;
                const uint32_t nt = out->scene.meshes[oit->second].n_triangles();
                if (instanced) {
                    for (const auto& ic : inst->arr) {
                        out->scene.instances.push_back(make_to_world(*ic));
                        const int32_t xi = (int32_t)out->scene.instances.size() - 1;
                        for (uint32_t t = 0; t < nt; ++t) out->scene.prims.push_back({SHAPE_TRIANGLE, oit->second, t, xi, mit->second});
                    }
                } else {
                    for (uint32_t t = 0; t < nt; ++t) out->scene.prims.push_back({SHAPE_TRIANGLE, oit->second, t, -1, mit->second});
                }
            }
        }
    // ---- make_all_lights (renderprocess.rs:921-966): `lights` and `infinite_lights` go through the same make_light ----
    auto resolve_path = [&](const std::string& file) { return file.size() && file[0] == '/' ? file : dir + "/" + file; };
    for (const std::string& f : tex.image_files) out->image_paths.push_back(resolve_path(f));
    auto make_light = [&](const json::ValuePtr& lc) -> rrt_light {
            rrt_light l;
            std::memset(&l, 0, sizeof(l));
            const std::string type = read_string(*lc, "light_type", "");
            const Transform xf = make_to_world(*lc);
            std::memcpy(l.to_world, xf.m.m, sizeof(l.to_world));
            if (type == "point") {
                l.kind = RRT_LIGHT_POINT;
                read_spectrum(*lc, "spectrum", l.intensity, 1.0);
            } else if (type == "distant") {
                l.kind = RRT_LIGHT_DISTANT;
                double li[3], sc[3], from[3], to[3];
                read_spectrum(*lc, "l", li, 1.0);
                read_spectrum(*lc, "scale", sc, 1.0);
                read_xyz(*lc, "from", from, 0, 0, 0);
                read_xyz(*lc, "to", to, 0, 0, 1);
                for (int k = 0; k < 3; ++k) {
                    l.intensity[k] = li[k] * sc[k];
                    l.dir[k] = from[k] - to[k];
                }
            } else if (type == "diffuse") {  // renderprocess.rs:999-1017
                l.kind = RRT_LIGHT_DIFFUSE_AREA;
                read_spectrum(*lc, "spectrum", l.intensity, 1.0);
                const Value* sc = lc->get("light_shape");
                if (!sc) throw std::runtime_error("Shape Required for a DiffuseLight! (renderprocess.rs:1015)");
                const std::string st = read_string(*sc, "shape_type", "");
                if (st == "sphere") {  // make_sphere (renderprocess.rs:1097-1106)
                    l.shape_kind = RRT_LIGHT_SHAPE_SPHERE;
                    const Transform sx = make_to_world(*sc);
                    std::memcpy(l.shape_to_world, sx.m.m, sizeof(l.shape_to_world));
                    std::memcpy(l.shape_to_world_inv, sx.inv.m, sizeof(l.shape_to_world_inv));
                    l.radius = read_f64(*sc, "radius", 1.0);
                    l.z_min = read_f64(*sc, "z_min", -l.radius);
                    l.z_max = read_f64(*sc, "z_max", l.radius);
                    l.phi_max_deg = read_f64(*sc, "phi_max", 360.0);
                } else if (st == "triangle") {  // mesh[tri_num] of a loaded obj (renderprocess.rs:1084-1090)
                    l.shape_kind = RRT_LIGHT_SHAPE_TRIANGLE;
                    auto oit = mesh_index.find(read_string(*sc, "obj_name", ""));
                    if (oit == mesh_index.end()) throw std::runtime_error("area light: unknown obj_name");
                    const TriangleMesh& m = out->scene.meshes[oit->second];
                    const uint32_t k = (uint32_t)read_f64(*sc, "tri_num", 0.0);
                    if (k >= m.n_triangles()) throw std::runtime_error("area light: tri_num out of range");
                    for (int v = 0; v < 3; ++v)
                        for (int c = 0; c < 3; ++c) l.tri_p[3 * v + c] = m.p[3 * (size_t)m.vi[3 * k + v] + c];
                    if (!m.n.empty() && !m.ni.empty()) {
                        l.tri_has_n = 1;
                        for (int v = 0; v < 3; ++v)
                            for (int c = 0; c < 3; ++c) l.tri_n[3 * v + c] = m.n[3 * (size_t)m.ni[3 * k + v] + c];
                    }
                } else {
                    throw std::runtime_error("Failed to parse a Shape (renderprocess.rs:1094)");
                }
            } else if (type == "infinite") {  // renderprocess.rs:1032-1046
                l.kind = RRT_LIGHT_INFINITE;
                double li[3], sc[3];
                read_spectrum(*lc, "l", li, 1.0);
                read_spectrum(*lc, "scale", sc, 1.0);
                for (int k = 0; k < 3; ++k) l.intensity[k] = li[k] * sc[k];  // carried; the reference never uses it (Q34)
                std::memcpy(l.shape_to_world_inv, xf.inv.m, sizeof(l.shape_to_world_inv));  // world_to_light
                l.env_image = (uint32_t)out->image_paths.size();
                out->image_paths.push_back(resolve_path(read_string(*lc, "mapname", "")));
            } else {
                throw std::runtime_error("light type '" + type + "' is outside the hot-path scope (point, distant, diffuse, infinite)");
            }
            return l;
    };
    if (const Value* a = root->get("lights"); a && a->is_array())
        for (const auto& lc : a->arr) out->lights.push_back(make_light(lc));
    if (const Value* a = root->get("infinite_lights"); a && a->is_array())
        for (const auto& lc : a->arr) out->infinite_lights.push_back(make_light(lc));

    // ---- make_integrator: film, camera, sampler, integrator ----
    const Value *ic = root->get("Integrator"), *sc = root->get("Sampler"), *fc = root->get("Film"), *cc = root->get("Camera");
    if (!ic || !sc || !fc || !cc) throw std::runtime_error("Failed to create Integrator: Integrator / Sampler / Film / Camera missing");
    rrt_render_desc& d = out->desc;
    std::memset(&d, 0, sizeof(d));
    d.xres = read_i64(*fc, "xres", 1280);
    d.yres = read_i64(*fc, "yres", 720);
    d.scale = read_f64(*fc, "scale", 1.0);
    d.diagonal_mm = read_f64(*fc, "diagonal", 35.0);
    d.max_sample_luminance = read_f64(*fc, "max_sample_luminance", INFINITY);
    const Value* flt = fc->get("Filter");
    if (!flt) throw std::runtime_error("Failed to create Film: filter_config not found");
    {
        const std::string ft = read_string(*flt, "filter_type", "BoxFilter");
        double def = 0.5;
        d.filter_kind = RRT_FILTER_BOX;
        if (ft == "TriangleFilter") {
            d.filter_kind = RRT_FILTER_TRIANGLE;
            def = 2.0;
        } else if (ft == "GaussianFilter") {
            d.filter_kind = RRT_FILTER_GAUSSIAN;
            def = 2.0;
        }
        d.filter_radius[0] = d.filter_radius[1] = def;
        if (const Value* r = flt->get("radius"); r && r->is_array() && r->arr.size() >= 2) {
            if (!r->arr[0]->is_number() || !r->arr[1]->is_number()) throw std::runtime_error("Film.Filter.radius entries must be numbers");
            d.filter_radius[0] = r->arr[0]->num;
            d.filter_radius[1] = r->arr[1]->num;
            if (!(d.filter_radius[0] > 0.0) || !(d.filter_radius[1] > 0.0) || !std::isfinite(d.filter_radius[0]) || !std::isfinite(d.filter_radius[1]))
                throw std::runtime_error("Film.Filter.radius must be finite and > 0");
        }
        d.filter_alpha = read_f64(*flt, "alpha", 2.0);
    }
    read_xyz(*cc, "world_pos", d.cam_pos, 0, 0, 0);
    read_xyz(*cc, "look", d.cam_look, 1, 1, 1);
    read_xyz(*cc, "up", d.cam_up, 0, 0, 1);
    d.shutter_open = read_f64(*cc, "shutter_open", 0.0);
    d.shutter_close = read_f64(*cc, "shutter_close", 1.0);
    d.aperture_diameter = read_f64(*cc, "aperture_diameter", 1.0);
    d.focus_distance = read_f64(*cc, "focus_distance", 10.0);
    d.simple_weighting = read_bool(*cc, "simple_weighting", true) ? 1 : 0;
    const Value* lens = cc->get("lens_data");
    if (!lens || !lens->is_array()) throw std::runtime_error("Camera.lens_data is required (renderprocess.rs:1379)");
    for (const auto& x : lens->arr) {
        if (!x->is_number()) throw std::runtime_error("Camera.lens_data entries must be numbers");
        out->lens_data.push_back(x->num);
    }
    d.lens_data = out->lens_data.data();
    d.n_lens_values = (uint32_t)out->lens_data.size();
    auto checked_u32 = [](int64_t v, int64_t lo, int64_t hi, const char* what) -> uint32_t {
        if (v < lo || v > hi) throw std::runtime_error(std::string(what) + " is out of range");
        return (uint32_t)v;
    };
    const std::string st = read_string(*sc, "sampler_type", "");
    d.seed = seed;
    if (st == "HaltonSampler") {
        const int64_t nsamp = read_i64(*sc, "nsamp", 16);
        if (nsamp < 1 || nsamp > (int64_t)1 << 32) throw std::runtime_error("Sampler.nsamp is out of range (1 .. 2^32)");
        d.nsamp = (uint64_t)nsamp;
        d.sample_at_center = read_bool(*sc, "sample_at_center", false) ? 1 : 0;
        d.sampler_kind = RRT_SAMPLER_HALTON;
    } else if (st == "StratifiedSampler") {  // renderprocess.rs:1308-1314
        d.sampler_kind = RRT_SAMPLER_STRATIFIED;
        d.strat_jitter = read_bool(*sc, "jitter", true) ? 1 : 0;
        d.strat_xsamp = checked_u32(read_i64(*sc, "xsamp", 4), 1, 256, "Sampler.xsamp");
        d.strat_ysamp = checked_u32(read_i64(*sc, "ysamp", 4), 1, 256, "Sampler.ysamp");
        d.strat_dimension = checked_u32(read_i64(*sc, "dimension", 4), 0, 255, "Sampler.dimension");
        d.nsamp = (uint64_t)d.strat_xsamp * d.strat_ysamp;
    } else {
        throw std::runtime_error("Unsupported Sampler type '" + st + "' (the reference panics: renderprocess.rs:1322)");
    }
    const std::string it = read_string(*ic, "integrator_type", "AO");
    if (it == "Path") {
        d.integrator_kind = RRT_INTEGRATOR_PATH;
        d.max_depth = checked_u32(read_i64(*ic, "max_depth", 5), 0, 64, "Integrator.max_depth");
        d.rr_threshold = read_f64(*ic, "rr_threshold", 1.0);
    } else if (it == "DirectLighting") {
        d.light_strategy = read_string(*ic, "light_strategy", "one") == "all" ? 1u : 0u;  // renderprocess.rs:1413-1417
        d.integrator_kind = RRT_INTEGRATOR_DIRECT;
        d.max_depth = checked_u32(read_i64(*ic, "max_depth", 5), 0, 64, "Integrator.max_depth");
        d.rr_threshold = 1.0;
    } else if (it == "Debug") {  // renderprocess.rs:1471-1481: max_depth only; always uniform_sample_all_lights
        d.integrator_kind = RRT_INTEGRATOR_DEBUG;
        d.light_strategy = 1u;
        d.max_depth = checked_u32(read_i64(*ic, "max_depth", 5), 0, 64, "Integrator.max_depth");
        d.rr_threshold = 1.0;
    } else {
        throw std::runtime_error("integrator '" + it + "' is outside the hot-path scope (Path, DirectLighting, Debug)");
    }
}

}  // namespace rrt
