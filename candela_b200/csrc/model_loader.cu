// Scene ingest without Assimp: a Wavefront OBJ reader that produces what
// ModelFileLoader.cpp:101-185 hands to the intersector — 32-byte Vertex records with the normal, tangent
// and UV packed as half floats exactly like glm::packHalf2x16 (ModelFileLoader.cpp:133-155), object-local
// indices with the per-mesh vertex offset already applied (BVHConstructor.cpp:981-1002) and one
// GlobalMeshNumber per triangle (ModelFileLoader.cpp:104-105: a running counter, one per mesh).
// Host code only.  The reference imports with aiProcess_JoinIdenticalVertices | Triangulate | CalcTangentSpace |
// GenUVCoords | FlipUVs | GenNormals (ModelFileLoader.cpp:243-252).  This reader does the same where the
// intersector can see it: one mesh per `usemtl` / `o` / `g` run (Assimp: one mesh per material), polygons
// fan-triangulated, v flipped (uv.y = 1 - uv.y), a flat face normal generated for corners that name no `vn`, and
// vertices joined per mesh when their position index, UV index, normal and tangent agree.  Tangents follow Assimp's
// CalcTangentsProcess (the step aiProcess_CalcTangentSpace names; Assimp itself is not in the reference tree — only its headers
// and a Windows import library — so the step is restated from its published algorithm and NOT pinned against a run of it):
// per face the UV-gradient tangent / bitangent, per corner Gram-Schmidt against the corner's normal, then one average per
// group of corners that share a position and a normal (dot >= 0.9999) and whose tangents and bitangents lie within 45 degrees.
// Nothing on the ray path reads them (GetData takes normal and UV; the tangent feeds the raster pass only).  Materials are
// recorded per mesh the way LoadMaterialTextures does (ModelFileLoader.cpp:31-99): albedo / normal texture path and ModelColor.
// The ORDER of the joined vertices is not Assimp's — which cannot matter: the builder and the traversal only ever see the
// arrays this loader returns.
#include <array>
#include <cerrno>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/candela_b200.h"

namespace {

// glm 0.9.8.5 detail::toFloat16 (glm/detail/type_half.inl:108-243): round to nearest with ties rounded UP
// in magnitude (bit 12 decides alone), denormals by shifting, NaN keeps its top 10 mantissa bits.
std::uint16_t to_float16(float f) {
    std::int32_t i;
    std::memcpy(&i, &f, 4);
    const int s = (i >> 16) & 0x00008000;
    int e = ((i >> 23) & 0x000000ff) - (127 - 15);
    int m = i & 0x007fffff;
    if (e <= 0) {
        if (e < -10) return (std::uint16_t)s;
        m = (m | 0x00800000) >> (1 - e);
        if (m & 0x00001000) m += 0x00002000;
        return (std::uint16_t)(s | (m >> 13));
    } else if (e == 0xff - (127 - 15)) {
        if (m == 0) return (std::uint16_t)(s | 0x7c00);
        m >>= 13;
        return (std::uint16_t)(s | 0x7c00 | m | (m == 0));
    }
    if (m & 0x00001000) {
        m += 0x00002000;
        if (m & 0x00800000) { m = 0; e += 1; }
    }
    if (e > 30) return (std::uint16_t)(s | 0x7c00);
    return (std::uint16_t)(s | (e << 10) | (m >> 13));
}

// One whole text line of any length (fgets into a buffer that grows): a statement is never cut in two — a polygon with thousands
// of corners, or a long comment whose tail would otherwise be read as a statement of its own.
bool read_line(std::FILE* f, std::vector<char>& buf) {
    size_t have = 0;
    for (;;) {
        if (!std::fgets(buf.data() + have, (int)(buf.size() - have), f)) return have > 0;
        have += std::strlen(buf.data() + have);
        if (have && buf[have - 1] == '\n') return true;
        if (have + 1 < buf.size()) return true;          // end of file without a newline (or an embedded NUL: the rest of the line is dropped)
        buf.resize(buf.size() * 2);
    }
}

struct V3 {
    float x = 0.0f, y = 0.0f, z = 0.0f;
    V3 operator-(const V3& o) const { return {x - o.x, y - o.y, z - o.z}; }
    V3 operator+(const V3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    V3 operator*(float f) const { return {x * f, y * f, z * f}; }
    float dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
    V3 cross(const V3& o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
    void normalize_safe() { const float l = std::sqrt(dot(*this)); if (l > 0.0f) { x /= l; y /= l; z /= l; } }
    void normalize() { const float l = std::sqrt(dot(*this)); x /= l; y /= l; z /= l; }
    bool special() const { return !std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z); }
};

// Assimp's CalcTangentsProcess::ProcessMesh restated (see the file header: unpinned).  `faces` holds vertex index triples; a vertex
// shared by several faces keeps the tangent of the last face that touches it before the smoothing pass, as in the original.
// Positions are matched bit for bit (the original: within 1e-4 of the mesh's bounding-box diagonal).
void calc_tangents(const std::vector<V3>& pos, const std::vector<V3>& nrm, const std::vector<std::array<float, 2>>& uv,
                   const std::vector<std::uint32_t>& faces, std::vector<V3>& tang) {
    const size_t n = pos.size();
    std::vector<V3> bitang(n);
    tang.assign(n, V3{});
    for (size_t f = 0; f + 2 < faces.size(); f += 3) {
        const std::uint32_t p0 = faces[f], p1 = faces[f + 1], p2 = faces[f + 2];
        const V3 v = pos[p1] - pos[p0], w = pos[p2] - pos[p0];
        float sx = uv[p1][0] - uv[p0][0], sy = uv[p1][1] - uv[p0][1];
        float tx = uv[p2][0] - uv[p0][0], ty = uv[p2][1] - uv[p0][1];
        const float dir = (tx * sy - ty * sx) < 0.0f ? -1.0f : 1.0f;
        if (sx * ty == sy * tx) { sx = 0.0f; sy = 1.0f; tx = 1.0f; ty = 0.0f; }  // degenerate UVs: a default mapping
        const V3 tangent{(w.x * sy - v.x * ty) * dir, (w.y * sy - v.y * ty) * dir, (w.z * sy - v.z * ty) * dir};
        const V3 bitangent{(w.x * sx - v.x * tx) * dir, (w.y * sx - v.y * tx) * dir, (w.z * sx - v.z * tx) * dir};
        for (const std::uint32_t p : {p0, p1, p2}) {
            V3 lt = tangent - nrm[p] * tangent.dot(nrm[p]);
            V3 lb = bitangent - nrm[p] * bitangent.dot(nrm[p]) - lt * bitangent.dot(lt);
            lt.normalize_safe();
            lb.normalize_safe();
            const bool bad_t = lt.special(), bad_b = lb.special();
            if (bad_t != bad_b) {
                if (bad_t) { lt = nrm[p].cross(lb); lt.normalize_safe(); }
                else { lb = lt.cross(nrm[p]); lb.normalize_safe(); }
            }
            tang[p] = lt;
            bitang[p] = lb;
        }
    }
    // smoothing: greedy groups seeded in vertex order
    struct PosKey { std::uint32_t b[3]; bool operator<(const PosKey& o) const { return std::memcmp(b, o.b, sizeof(b)) < 0; } };
    std::map<PosKey, std::vector<std::uint32_t>> at;
    auto key_of = [&](const V3& p) {
        const float c[3] = {p.x + 0.0f, p.y + 0.0f, p.z + 0.0f};  // -0 -> +0
        PosKey k;
        std::memcpy(k.b, c, sizeof(c));
        return k;
    };
    for (size_t a = 0; a < n; ++a) at[key_of(pos[a])].push_back((std::uint32_t)a);
    const float limit = std::cos(45.0f * 3.14159265358979323846f / 180.0f), same_normal = 0.9999f;
    std::vector<char> done(n, 0);
    std::vector<std::uint32_t> close;
    for (size_t a = 0; a < n; ++a) {
        if (done[a]) continue;
        close.clear();
        for (const std::uint32_t b : at[key_of(pos[a])]) {
            if (done[b]) continue;
            if (b != a && (nrm[b].dot(nrm[a]) < same_normal || tang[b].dot(tang[a]) < limit || bitang[b].dot(bitang[a]) < limit)) continue;
            close.push_back(b);
            done[b] = 1;
        }
        if (close.size() < 2) continue;
        V3 st, sb;
        for (const std::uint32_t b : close) { st = st + tang[b]; sb = sb + bitang[b]; }
        st.normalize();
        sb.normalize();
        for (const std::uint32_t b : close) { tang[b] = st; bitang[b] = sb; }
    }
}

std::string parent_dir(const std::string& path) {  // std::filesystem::path::parent_path as LoadMaterialTextures uses it (:33-35)
    const size_t slash = path.find_last_of("/\\");
    return slash == std::string::npos ? std::string() : path.substr(0, slash);
}

}  // namespace

struct cndl_model {
    std::vector<cndl_vertex> vertices;
    std::vector<std::uint32_t> indices;   // object-local: mesh-local index + vertices of earlier meshes
    std::vector<std::int32_t> mesh_ids;   // one per triangle
    std::vector<std::uint32_t> mesh_first_vertex, mesh_first_index;
    std::vector<std::string> mesh_names;
    std::vector<std::string> mesh_albedo, mesh_normal;     // _MeshMaterialData::Albedo / Normal (ModelFileLoader.h:21-25)
    std::vector<std::array<float, 3>> mesh_color;          // _MeshMaterialData::ModelColor
};

extern "C" {

uint32_t cndl_pack_half2x16(float x, float y) { return (uint32_t)to_float16(x) | ((uint32_t)to_float16(y) << 16); }

int cndl_model_load_obj(const char* path, int32_t first_mesh_number, cndl_model** out, char* err, size_t err_cap) try {
    auto fail = [&](const std::string& msg) {
        if (err && err_cap) std::snprintf(err, err_cap, "%s", msg.c_str());
        return (int)CNDL_ERR_INVALID;
    };
    if (!path || !out) return fail("null argument");
    *out = nullptr;
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return fail(std::string("cannot open ") + path + ": " + std::strerror(errno));
    struct FileCloser { std::FILE* f; ~FileCloser() { if (f) std::fclose(f); } } closer{f};
    std::unique_ptr<cndl_model> owner(new cndl_model);
    cndl_model* M = owner.get();

    const std::string dir = parent_dir(path);
    std::vector<float> P, N, T;  // v (3), vn (3), vt (2)
    // One triangle corner as the file names it.  n >= 0: index into the file's `vn` list; n == -1: generated face normal, whose
    // value is in g[].  Assimp's OBJ importer emits one vertex per corner like this; tangents are computed on them, then they are joined.
    struct Corner {
        long v, t, n;
        float g[3];
    };
    struct Key {
        long v, t, n;
        std::uint32_t g[3], tan[3];
        bool operator==(const Key& o) const { return v == o.v && t == o.t && n == o.n && std::memcmp(g, o.g, sizeof(g)) == 0 && std::memcmp(tan, o.tan, sizeof(tan)) == 0; }
    };
    struct KeyHash {
        size_t operator()(const Key& k) const {
            return (size_t)k.v * 73856093u ^ (size_t)(k.t + 1) * 19349663u ^ (size_t)(k.n + 1) * 83492791u ^ k.g[0] ^ (k.g[1] * 31u) ^ (k.g[2] * 131u) ^
                   (k.tan[0] * 7u) ^ (k.tan[1] * 977u) ^ (k.tan[2] * 4099u);
        }
    };
    // materials of the `mtllib` files: Kd -> AI_MATKEY_COLOR_DIFFUSE, map_Kd -> aiTextureType_DIFFUSE, norm / map_Kn -> aiTextureType_NORMALS
    // (map_bump / bump are aiTextureType_HEIGHT in Assimp's OBJ importer and therefore NOT what ModelFileLoader.cpp:58 asks for)
    struct Mtl { float kd[3] = {0.6f, 0.6f, 0.6f}; std::string map_kd, norm; };
    std::map<std::string, Mtl> materials;
    auto read_mtl = [&](const std::string& name) {
        std::FILE* mf = std::fopen((dir.empty() ? name : dir + "/" + name).c_str(), "rb");
        if (!mf) return;  // Assimp logs and carries on with default materials
        std::vector<char> ml(1 << 12);
        Mtl* cur = nullptr;
        auto rest = [](char* e) {  // last blank-separated token of the statement: texture options (-bm 1.0 ...) precede the file name
            std::string r(e);
            while (!r.empty() && (r.back() == '\n' || r.back() == '\r' || r.back() == ' ' || r.back() == '\t')) r.pop_back();
            const size_t sp = r.find_last_of(" \t");
            return sp == std::string::npos ? r : r.substr(sp + 1);
        };
        while (read_line(mf, ml)) {
            char* q = ml.data();
            while (*q == ' ' || *q == '\t') ++q;
            if (std::strncmp(q, "newmtl", 6) == 0) cur = &materials[rest(q + 6)];
            else if (!cur) continue;
            else if (q[0] == 'K' && q[1] == 'd' && (q[2] == ' ' || q[2] == '\t')) { char* e = q + 2; for (int k = 0; k < 3; ++k) cur->kd[k] = std::strtof(e, &e); }
            else if (std::strncmp(q, "map_Kd", 6) == 0 && (q[6] == ' ' || q[6] == '\t')) cur->map_kd = rest(q + 6);
            else if (std::strncmp(q, "map_Kn", 6) == 0 && (q[6] == ' ' || q[6] == '\t')) cur->norm = rest(q + 6);
            else if (std::strncmp(q, "norm", 4) == 0 && (q[4] == ' ' || q[4] == '\t')) cur->norm = rest(q + 4);
        }
        std::fclose(mf);
    };

    bool mesh_open = false;
    std::string pending_name = "default", current_material, mesh_material;
    std::vector<Corner> mesh_corners;  // three per triangle of the open mesh
    auto close_mesh = [&]() {
        if (!mesh_open) return;
        mesh_open = false;
        const size_t nc = mesh_corners.size();
        std::vector<V3> tang(nc);
        bool has_uv = false;
        for (const Corner& c : mesh_corners) has_uv = has_uv || c.t >= 0;
        auto normal_of = [&](const Corner& c) { return c.n >= 0 ? V3{N[3 * c.n], N[3 * c.n + 1], N[3 * c.n + 2]} : V3{c.g[0], c.g[1], c.g[2]}; };
        auto uv_of = [&](const Corner& c) { return c.t >= 0 ? std::array<float, 2>{T[2 * c.t], 1.0f - T[2 * c.t + 1]} : std::array<float, 2>{0.0f, 0.0f}; };  // aiProcess_FlipUVs
        if (has_uv) {  // CalcTangentsProcess needs a UV channel; without one mTangents stays null and the reference packs zeros (:138-143)
            std::vector<V3> pos(nc), nrm(nc);
            std::vector<std::array<float, 2>> uv(nc);
            std::vector<std::uint32_t> faces(nc);
            for (size_t k = 0; k < nc; ++k) {
                const Corner& c = mesh_corners[k];
                pos[k] = V3{P[3 * c.v], P[3 * c.v + 1], P[3 * c.v + 2]};
                nrm[k] = normal_of(c);
                uv[k] = uv_of(c);
                faces[k] = (std::uint32_t)k;
            }
            calc_tangents(pos, nrm, uv, faces, tang);
        }
        std::unordered_map<Key, std::uint32_t, KeyHash> seen;  // aiProcess_JoinIdenticalVertices, per mesh
        for (size_t k = 0; k < nc; ++k) {
            const Corner& c = mesh_corners[k];
            Key key{c.v, c.t, c.n, {0, 0, 0}, {0, 0, 0}};
            std::memcpy(key.g, c.g, sizeof(key.g));
            const float tg[3] = {tang[k].x + 0.0f, tang[k].y + 0.0f, tang[k].z + 0.0f};
            std::memcpy(key.tan, tg, sizeof(tg));
            auto it = seen.find(key);
            if (it != seen.end()) { M->indices.push_back(it->second); continue; }
            cndl_vertex v;
            std::memset(&v, 0, sizeof(v));
            v.position[0] = P[3 * c.v]; v.position[1] = P[3 * c.v + 1]; v.position[2] = P[3 * c.v + 2]; v.position[3] = 1.0f;
            const V3 nv = normal_of(c);
            const std::array<float, 2> tuv = uv_of(c);
            v.normal_tangent[0] = cndl_pack_half2x16(nv.x, nv.y);          // data.x = packHalf2x16(vnormal.xy)         (ModelFileLoader.cpp:150)
            v.normal_tangent[1] = cndl_pack_half2x16(nv.z, tang[k].x);     // data.y = packHalf2x16(vnormal.z, vtan.x)  (:151)
            v.normal_tangent[2] = cndl_pack_half2x16(tang[k].y, tang[k].z);  // data.z = packHalf2x16(vtan.yz)           (:152)
            v.texcoords = cndl_pack_half2x16(tuv[0], tuv[1]);              // :133-136, (0,0) when the mesh has no UVs (:148)
            const std::uint32_t idx = (std::uint32_t)M->vertices.size();   // object-local = mesh-local + vertices of earlier meshes
            M->vertices.push_back(v);
            seen.emplace(key, idx);
            M->indices.push_back(idx);
        }
        mesh_corners.clear();
        // LoadMaterialTextures (:31-99): texture_path + "/" + name, "" when the material names no texture; ModelColor = diffuse colour
        const auto mt = materials.find(mesh_material);
        const Mtl m = mt == materials.end() ? Mtl{} : mt->second;
        M->mesh_albedo.push_back(m.map_kd.empty() ? std::string() : dir + "/" + m.map_kd);
        M->mesh_normal.push_back(m.norm.empty() ? std::string() : dir + "/" + m.norm);
        M->mesh_color.push_back({m.kd[0], m.kd[1], m.kd[2]});
    };
    auto open_mesh = [&]() {
        M->mesh_first_vertex.push_back((std::uint32_t)M->vertices.size());
        M->mesh_first_index.push_back((std::uint32_t)M->indices.size());
        M->mesh_names.push_back(pending_name);
        mesh_material = current_material;
        mesh_open = true;
    };

    std::vector<char> line(1 << 16);
    long lineno = 0;
    std::string problem;
    while (read_line(f, line)) {
        ++lineno;
        char* s = line.data();
        while (*s == ' ' || *s == '\t') ++s;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            char* e = s + 1;
            for (int k = 0; k < 3; ++k) P.push_back(std::strtof(e, &e));
        } else if (s[0] == 'v' && s[1] == 'n') {
            char* e = s + 2;
            for (int k = 0; k < 3; ++k) N.push_back(std::strtof(e, &e));
        } else if (s[0] == 'v' && s[1] == 't') {
            char* e = s + 2;
            for (int k = 0; k < 2; ++k) T.push_back(std::strtof(e, &e));
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            std::vector<Corner> corners;
            char* e = s + 1;
            while (true) {
                while (*e == ' ' || *e == '\t') ++e;
                if (*e == '\0' || *e == '\n' || *e == '\r' || *e == '#') break;
                Corner k{0, -1, -1, {0.0f, 0.0f, 0.0f}};
                long v = std::strtol(e, &e, 10), t = 0, n = 0;
                bool has_t = false, has_n = false;
                if (*e == '/') {
                    ++e;
                    if (*e != '/') { t = std::strtol(e, &e, 10); has_t = true; }
                    if (*e == '/') { ++e; n = std::strtol(e, &e, 10); has_n = true; }
                }
                const long nv = (long)P.size() / 3, nt = (long)T.size() / 2, nn = (long)N.size() / 3;
                k.v = v > 0 ? v - 1 : nv + v;
                k.t = has_t ? (t > 0 ? t - 1 : nt + t) : -1;
                k.n = has_n ? (n > 0 ? n - 1 : nn + n) : -1;
                if (k.v < 0 || k.v >= nv || (has_t && (k.t < 0 || k.t >= nt)) || (has_n && (k.n < 0 || k.n >= nn))) {
                    problem = "index out of range on line " + std::to_string(lineno);
                    break;
                }
                corners.push_back(k);
            }
            if (!problem.empty()) break;
            if (corners.size() < 3) continue;
            if (!mesh_open) open_mesh();
            {  // aiProcess_GenNormals: corners without a `vn` get the face's normal
                const float* a = &P[3 * corners[0].v];
                const float* b = &P[3 * corners[1].v];
                const float* c = &P[3 * corners[2].v];
                const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
                float fn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                const float len = std::sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
                if (len > 0.0f) { fn[0] /= len; fn[1] /= len; fn[2] /= len; }
                for (auto& k : corners)
                    if (k.n < 0) { k.g[0] = fn[0] + 0.0f; k.g[1] = fn[1] + 0.0f; k.g[2] = fn[2] + 0.0f; }
            }
            for (size_t c = 1; c + 1 < corners.size(); ++c) {  // triangle fan, like aiProcess_Triangulate on convex polygons
                mesh_corners.push_back(corners[0]);
                mesh_corners.push_back(corners[c]);
                mesh_corners.push_back(corners[c + 1]);
            }
        } else if (std::strncmp(s, "usemtl", 6) == 0 || ((s[0] == 'o' || s[0] == 'g') && (s[1] == ' ' || s[1] == '\t'))) {
            char* e = s + (s[0] == 'u' ? 6 : 1);
            while (*e == ' ' || *e == '\t') ++e;
            std::string name(e);
            while (!name.empty() && (name.back() == '\n' || name.back() == '\r' || name.back() == ' ')) name.pop_back();
            pending_name = name.empty() ? "default" : name;
            if (s[0] == 'u') current_material = name;
            close_mesh();  // a new mesh starts at the next face
        } else if (std::strncmp(s, "mtllib", 6) == 0 && (s[6] == ' ' || s[6] == '\t')) {
            char* e = s + 6;
            while (*e == ' ' || *e == '\t') ++e;
            std::string name(e);
            while (!name.empty() && (name.back() == '\n' || name.back() == '\r' || name.back() == ' ')) name.pop_back();
            if (!name.empty()) read_mtl(name);
        }
    }
    close_mesh();
    if (!problem.empty() || M->indices.empty()) return fail(problem.empty() ? std::string("no faces in ") + path : problem);
    // one GlobalMeshNumber per mesh, consecutive from first_mesh_number; one entry per triangle
    M->mesh_ids.resize(M->indices.size() / 3);
    for (size_t m = 0; m < M->mesh_first_index.size(); ++m) {
        const size_t lo = M->mesh_first_index[m] / 3, hi = (m + 1 < M->mesh_first_index.size() ? M->mesh_first_index[m + 1] : M->indices.size()) / 3;
        for (size_t t = lo; t < hi; ++t) M->mesh_ids[t] = first_mesh_number + (std::int32_t)m;
    }
    *out = owner.release();
    return CNDL_OK;
} catch (const std::bad_alloc&) {
    return CNDL_ERR_OOM;
} catch (...) {
    return CNDL_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------
// glTF 2.0 (.gltf + external / embedded buffers, .glb).  The reference walks Assimp's node tree depth first and
// adds every mesh a node references as it stands, WITHOUT the node's transform (ProcessAssimpNode,
// ModelFileLoader.cpp:187-227), one GlobalMeshNumber per mesh in that order; Assimp makes one mesh per glTF
// primitive.  Same here: scene -> nodes depth first -> mesh -> primitives.  FlipUVs applies (v = 1 - v); primitives
// without normals get flat face normals (GenNormals).  Sparse accessors and Draco are not supported.
}  // extern "C"

namespace {

struct JVal {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    double num = 0.0;
    bool b = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* find(const char* key) const {
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    long integer(const char* key, long dflt) const {
        const JVal* v = find(key);
        return (v && v->type == Num) ? (long)v->num : dflt;
    }
};

struct JParser {
    const char* p;
    const char* end;
    bool ok = true;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
    bool lit(const char* s) {
        const size_t n = std::strlen(s);
        if ((size_t)(end - p) >= n && std::memcmp(p, s, n) == 0) { p += n; return true; }
        return false;
    }
    std::string string() {
        std::string out;
        if (p >= end || *p != '"') { ok = false; return out; }
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {  // BMP code point -> UTF-8 (names only; never on the data path)
                        unsigned cp = 0;
                        for (int k = 0; k < 4 && p + 1 < end; ++k) { ++p; const char c = *p; cp = cp * 16 + (unsigned)(c <= '9' ? c - '0' : (c | 32) - 'a' + 10); }
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += *p;
                }
                ++p;
            } else {
                out += *p++;
            }
        }
        if (p >= end) { ok = false; return out; }
        ++p;
        return out;
    }
    JVal value(int depth = 0) {
        JVal v;
        ws();
        if (p >= end || depth > 64) { ok = false; return v; }
        if (*p == '{') {
            v.type = JVal::Obj;
            ++p;
            ws();
            if (p < end && *p == '}') { ++p; return v; }
            while (ok) {
                ws();
                std::string k = string();
                ws();
                if (p >= end || *p != ':') { ok = false; break; }
                ++p;
                v.obj.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                ok = false;
            }
        } else if (*p == '[') {
            v.type = JVal::Arr;
            ++p;
            ws();
            if (p < end && *p == ']') { ++p; return v; }
            while (ok) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                ok = false;
            }
        } else if (*p == '"') {
            v.type = JVal::Str;
            v.str = string();
        } else if (lit("true")) { v.type = JVal::Bool; v.b = true; }
        else if (lit("false")) { v.type = JVal::Bool; }
        else if (lit("null")) {}
        else {
            char* e = nullptr;
            v.type = JVal::Num;
            v.num = std::strtod(p, &e);
            if (e == p) ok = false;
            p = e;
        }
        return v;
    }
};

bool read_file(const std::string& path, std::vector<unsigned char>& out) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && std::fread(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

bool base64_decode(const char* s, size_t n, std::vector<unsigned char>& out) {
    unsigned acc = 0;
    int bits = 0;
    for (size_t i = 0; i < n; ++i) {
        const char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else if (c == '=' || c == '\n' || c == '\r') continue;
        else return false;
        acc = (acc << 6) | (unsigned)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((unsigned char)((acc >> bits) & 0xFF)); }
    }
    return true;
}

struct Accessor {
    const unsigned char* base = nullptr;  // first element
    size_t stride = 0, count = 0;
    int component = 0, width = 0;         // componentType, components per element
    bool normalized = false;
    float get(size_t i, int c) const {
        const unsigned char* e = base + i * stride;
        switch (component) {
            case 5126: { float f; std::memcpy(&f, e + 4 * c, 4); return f; }
            case 5121: { const float v = (float)e[c]; return normalized ? v / 255.0f : v; }
            case 5123: { std::uint16_t u; std::memcpy(&u, e + 2 * c, 2); return normalized ? (float)u / 65535.0f : (float)u; }
            case 5120: { const float v = (float)(signed char)e[c]; return normalized ? std::fmax(v / 127.0f, -1.0f) : v; }
            case 5122: { std::int16_t u; std::memcpy(&u, e + 2 * c, 2); return normalized ? std::fmax((float)u / 32767.0f, -1.0f) : (float)u; }
            case 5125: { std::uint32_t u; std::memcpy(&u, e + 4 * c, 4); return (float)u; }
        }
        return 0.0f;
    }
    std::uint32_t index(size_t i) const {
        const unsigned char* e = base + i * stride;
        switch (component) {
            case 5121: return e[0];
            case 5123: { std::uint16_t u; std::memcpy(&u, e, 2); return u; }
            case 5125: { std::uint32_t u; std::memcpy(&u, e, 4); return u; }
        }
        return 0;
    }
};

int component_size(int t) { return (t == 5120 || t == 5121) ? 1 : (t == 5122 || t == 5123) ? 2 : (t == 5125 || t == 5126) ? 4 : 0; }

}  // namespace

extern "C" {

int cndl_model_load_gltf(const char* path, int32_t first_mesh_number, cndl_model** out, char* err, size_t err_cap) try {
    auto fail = [&](const std::string& msg) {
        if (err && err_cap) std::snprintf(err, err_cap, "%s", msg.c_str());
        return (int)CNDL_ERR_INVALID;
    };
    if (!path || !out) return fail("null argument");
    *out = nullptr;
    std::vector<unsigned char> file;
    if (!read_file(path, file)) return fail(std::string("cannot open ") + path);
    const std::string spath(path);
    const size_t slash = spath.find_last_of("/\\");
    const std::string dir = slash == std::string::npos ? std::string() : spath.substr(0, slash + 1);

    std::vector<unsigned char> glb_bin;
    const char* json_begin = reinterpret_cast<const char*>(file.data());
    size_t json_len = file.size();
    if (file.size() >= 20 && std::memcmp(file.data(), "glTF", 4) == 0) {  // .glb: header, JSON chunk, optional BIN chunk
        std::uint32_t len0, type0;
        std::memcpy(&len0, file.data() + 12, 4);
        std::memcpy(&type0, file.data() + 16, 4);
        if (type0 != 0x4E4F534Au || 20ull + len0 > file.size()) return fail("malformed .glb");
        json_begin = reinterpret_cast<const char*>(file.data() + 20);
        json_len = len0;
        const size_t at = 20 + (size_t)len0;
        if (at + 8 <= file.size()) {
            std::uint32_t len1, type1;
            std::memcpy(&len1, file.data() + at, 4);
            std::memcpy(&type1, file.data() + at + 4, 4);
            if (type1 == 0x004E4942u && (size_t)len1 <= file.size() - (at + 8)) glb_bin.assign(file.data() + at + 8, file.data() + at + 8 + len1);
        }
    }
    JParser jp{json_begin, json_begin + json_len};
    const JVal root = jp.value();
    if (!jp.ok || root.type != JVal::Obj) return fail(std::string(path) + ": not valid JSON");
    const JVal *jbuffers = root.find("buffers"), *jviews = root.find("bufferViews"), *jacc = root.find("accessors"), *jmeshes = root.find("meshes"),
               *jnodes = root.find("nodes"), *jscenes = root.find("scenes"), *jmats = root.find("materials"), *jtex = root.find("textures"),
               *jimages = root.find("images");
    const std::string texture_dir = parent_dir(spath);
    // textures[i].source -> images[j].uri; an image held in a bufferView is named "*j", Assimp's name for an embedded texture
    auto texture_name = [&](const JVal* slot) -> std::string {
        if (!slot || slot->type != JVal::Obj || !jtex) return std::string();
        const long t = slot->integer("index", -1);
        if (t < 0 || (size_t)t >= jtex->arr.size()) return std::string();
        const long img = jtex->arr[(size_t)t].integer("source", -1);
        if (!jimages || img < 0 || (size_t)img >= jimages->arr.size()) return std::string();
        const JVal* uri = jimages->arr[(size_t)img].find("uri");
        if (uri && uri->type == JVal::Str && uri->str.compare(0, 5, "data:") != 0) return uri->str;
        return "*" + std::to_string(img);
    };
    if (!jviews || !jacc || !jmeshes || !jbuffers) return fail(std::string(path) + ": no meshes / accessors / bufferViews / buffers");

    std::vector<std::vector<unsigned char>> buffers(jbuffers->arr.size());
    for (size_t i = 0; i < buffers.size(); ++i) {
        const JVal* uri = jbuffers->arr[i].find("uri");
        if (!uri) { buffers[i] = glb_bin; continue; }
        const std::string& u = uri->str;
        if (u.compare(0, 5, "data:") == 0) {
            const size_t comma = u.find(',');
            if (comma == std::string::npos || !base64_decode(u.c_str() + comma + 1, u.size() - comma - 1, buffers[i])) return fail("bad data: URI in buffer");
        } else if (!read_file(dir + u, buffers[i])) {
            return fail("cannot open buffer " + dir + u);
        }
    }
    auto accessor = [&](long idx, Accessor& a, std::string& why) -> bool {
        if (idx < 0 || (size_t)idx >= jacc->arr.size()) { why = "accessor index out of range"; return false; }
        const JVal& j = jacc->arr[idx];
        if (j.find("sparse")) { why = "sparse accessors are not supported"; return false; }
        const long view = j.integer("bufferView", -1);
        if (view < 0 || (size_t)view >= jviews->arr.size()) { why = "accessor without a bufferView"; return false; }
        const JVal& v = jviews->arr[view];
        const long buf = v.integer("buffer", -1);
        if (buf < 0 || (size_t)buf >= buffers.size()) { why = "bufferView names no buffer"; return false; }
        const JVal* type = j.find("type");
        a.component = (int)j.integer("componentType", 0);
        a.width = !type ? 0 : type->str == "SCALAR" ? 1 : type->str == "VEC2" ? 2 : type->str == "VEC3" ? 3 : type->str == "VEC4" ? 4 : 0;
        const long jcount = j.integer("count", 0);
        if (jcount < 0) { why = "negative accessor count"; return false; }
        a.count = (size_t)jcount;
        const JVal* nrm = j.find("normalized");
        a.normalized = nrm && nrm->type == JVal::Bool && nrm->b;
        const int cs = component_size(a.component);
        if (!cs || !a.width) { why = "unsupported accessor type"; return false; }
        const size_t elem = (size_t)cs * a.width;
        const long bs = v.integer("byteStride", 0), off_v = v.integer("byteOffset", 0), off_a = j.integer("byteOffset", 0);
        if (bs < 0 || off_v < 0 || off_a < 0) { why = "negative byteStride / byteOffset"; return false; }
        a.stride = bs > 0 ? (size_t)bs : elem;
        // overflow-safe range check: off + (count - 1) * stride + elem <= size, without forming a product that can wrap
        const size_t size = buffers[buf].size(), off = (size_t)off_v + (size_t)off_a;
        if (off < (size_t)off_v || off > size || elem > size - off) { if (a.count) { why = "accessor runs past its buffer"; return false; } }
        else if (a.count && a.count - 1 > (size - off - elem) / a.stride) { why = "accessor runs past its buffer"; return false; }
        a.base = buffers[buf].data() + off;
        return true;
    };

    std::unique_ptr<cndl_model> owner(new cndl_model);
    cndl_model* M = owner.get();
    std::string why;
    auto add_mesh = [&](const JVal& mesh) -> bool {
        const JVal* prims = mesh.find("primitives");
        const JVal* name = mesh.find("name");
        if (!prims) return true;
        for (const JVal& prim : prims->arr) {
            const long mode = prim.integer("mode", 4);
            if (mode < 4 || mode > 6) continue;  // points / lines are not triangles
            const JVal* attrs = prim.find("attributes");
            if (!attrs) continue;
            Accessor pos, nor, uv, idx, tan;
            if (!accessor(attrs->integer("POSITION", -1), pos, why) || pos.width != 3) { if (why.empty()) why = "POSITION must be VEC3"; return false; }
            const bool has_n = attrs->find("NORMAL") != nullptr, has_uv = attrs->find("TEXCOORD_0") != nullptr, has_i = prim.find("indices") != nullptr;
            if (has_n && (!accessor(attrs->integer("NORMAL", -1), nor, why) || nor.count < pos.count)) { if (why.empty()) why = "NORMAL shorter than POSITION"; return false; }
            if (has_uv && (!accessor(attrs->integer("TEXCOORD_0", -1), uv, why) || uv.count < pos.count)) { if (why.empty()) why = "TEXCOORD_0 shorter than POSITION"; return false; }
            if (has_i && !accessor(prim.integer("indices", -1), idx, why)) return false;
            // a TANGENT attribute is imported as it stands (CalcTangentsProcess skips meshes that already carry tangents)
            const bool has_t = has_n && attrs->find("TANGENT") != nullptr && accessor(attrs->integer("TANGENT", -1), tan, why) && tan.width >= 3 && tan.count >= pos.count;
            why.clear();
            // corner list in triangle order
            const size_t n_src = has_i ? idx.count : pos.count;
            std::vector<std::uint32_t> corners;
            auto src = [&](size_t k) { return has_i ? idx.index(k) : (std::uint32_t)k; };
            if (mode == 4) { for (size_t k = 0; k + 2 < n_src + 0 && k + 2 < n_src; k += 3) { corners.push_back(src(k)); corners.push_back(src(k + 1)); corners.push_back(src(k + 2)); } }
            else if (mode == 5) { for (size_t k = 0; k + 2 < n_src; ++k) { corners.push_back(src(k + (k & 1))); corners.push_back(src(k + 1 - (k & 1))); corners.push_back(src(k + 2)); } }
            else { for (size_t k = 1; k + 1 < n_src; ++k) { corners.push_back(src(0)); corners.push_back(src(k)); corners.push_back(src(k + 1)); } }
            for (std::uint32_t c : corners)
                if (c >= pos.count) { why = "index past the vertex count"; return false; }
            if (corners.empty()) continue;
            const std::uint32_t first_vertex = (std::uint32_t)M->vertices.size();
            M->mesh_first_vertex.push_back(first_vertex);
            M->mesh_first_index.push_back((std::uint32_t)M->indices.size());
            M->mesh_names.push_back(name && name->type == JVal::Str ? name->str : std::string("mesh"));
            {   // LoadMaterialTextures (ModelFileLoader.cpp:31-99): base colour texture, normal texture, diffuse colour
                std::string albedo, normal;
                std::array<float, 3> color{1.0f, 1.0f, 1.0f};  // glTF default baseColorFactor, also Assimp's default glTF material
                const long mat = prim.integer("material", -1);
                if (jmats && mat >= 0 && (size_t)mat < jmats->arr.size()) {
                    const JVal& jm = jmats->arr[(size_t)mat];
                    if (const JVal* pbr = jm.find("pbrMetallicRoughness")) {
                        if (const JVal* bc = pbr->find("baseColorFactor"))
                            for (size_t k = 0; k < 3 && k < bc->arr.size(); ++k) color[k] = (float)bc->arr[k].num;
                        albedo = texture_name(pbr->find("baseColorTexture"));
                    }
                    normal = texture_name(jm.find("normalTexture"));
                }
                M->mesh_albedo.push_back(albedo.empty() ? std::string() : texture_dir + "/" + albedo);
                M->mesh_normal.push_back(normal.empty() ? std::string() : texture_dir + "/" + normal);
                M->mesh_color.push_back(color);
            }
            auto make_vertex = [&](std::uint32_t v, const float* n) {
                cndl_vertex o;
                std::memset(&o, 0, sizeof(o));
                o.position[0] = pos.get(v, 0); o.position[1] = pos.get(v, 1); o.position[2] = pos.get(v, 2); o.position[3] = 1.0f;
                const float tu = has_uv ? uv.get(v, 0) : 0.0f, tv = has_uv ? 1.0f - uv.get(v, 1) : 0.0f;  // aiProcess_FlipUVs
                o.normal_tangent[0] = cndl_pack_half2x16(n[0], n[1]);
                o.normal_tangent[1] = cndl_pack_half2x16(n[2], 0.0f);
                o.normal_tangent[2] = cndl_pack_half2x16(0.0f, 0.0f);
                o.texcoords = cndl_pack_half2x16(tu, tv);
                return o;
            };
            std::vector<V3> vertex_normals;
            std::vector<std::array<float, 2>> vertex_uvs;
            // tangents of the mesh's vertices [first_vertex, end) over faces given as mesh-local index triples (ModelFileLoader.cpp:138-152)
            auto add_tangents = [&](const std::vector<std::uint32_t>& faces) {
                if (!has_uv) return;  // no UV channel: mTangents stays null, the reference packs zeros
                const size_t nv = M->vertices.size() - first_vertex;
                std::vector<V3> tg(nv);
                if (has_t) {
                    for (size_t v = 0; v < nv; ++v) tg[v] = V3{tan.get(v, 0), tan.get(v, 1), tan.get(v, 2)};
                } else {
                    std::vector<V3> p(nv), nn(nv);
                    std::vector<std::array<float, 2>> tuv(nv);
                    for (size_t v = 0; v < nv; ++v) {
                        const cndl_vertex& o = M->vertices[first_vertex + v];
                        p[v] = V3{o.position[0], o.position[1], o.position[2]};
                        nn[v] = vertex_normals[v];
                        tuv[v] = vertex_uvs[v];
                    }
                    calc_tangents(p, nn, tuv, faces, tg);
                }
                for (size_t v = 0; v < nv; ++v) {
                    cndl_vertex& o = M->vertices[first_vertex + v];
                    o.normal_tangent[1] = cndl_pack_half2x16(vertex_normals[v].z, tg[v].x);
                    o.normal_tangent[2] = cndl_pack_half2x16(tg[v].y, tg[v].z);
                }
            };
            auto remember = [&](std::uint32_t v, const float* n) {
                vertex_normals.push_back(V3{n[0], n[1], n[2]});
                vertex_uvs.push_back({has_uv ? uv.get(v, 0) : 0.0f, has_uv ? 1.0f - uv.get(v, 1) : 0.0f});
            };
            if (has_n) {  // indexed as in the file
                for (std::uint32_t v = 0; v < pos.count; ++v) {
                    const float n[3] = {nor.get(v, 0), nor.get(v, 1), nor.get(v, 2)};
                    M->vertices.push_back(make_vertex(v, n));
                    remember(v, n);
                }
                for (std::uint32_t c : corners) M->indices.push_back(first_vertex + c);
                add_tangents(corners);
            } else {  // aiProcess_GenNormals: flat shading, one vertex per corner
                std::vector<std::uint32_t> faces;
                for (size_t k = 0; k + 2 < corners.size() + 0 && k + 2 < corners.size(); k += 3) {
                    float p[3][3];
                    for (int c = 0; c < 3; ++c)
                        for (int d = 0; d < 3; ++d) p[c][d] = pos.get(corners[k + c], d);
                    const float e1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]}, e2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
                    float fn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                    const float len = std::sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
                    if (len > 0.0f) { fn[0] /= len; fn[1] /= len; fn[2] /= len; }
                    for (int c = 0; c < 3; ++c) {
                        faces.push_back((std::uint32_t)(M->vertices.size() - first_vertex));
                        M->indices.push_back((std::uint32_t)M->vertices.size());
                        M->vertices.push_back(make_vertex(corners[k + c], fn));
                        remember(corners[k + c], fn);
                    }
                }
                add_tangents(faces);
            }
        }
        return true;
    };
    // scene -> nodes, depth first (ProcessAssimpNode: a node's meshes, then its children)
    std::vector<long> stack;
    if (jscenes && jnodes && !jscenes->arr.empty()) {
        const long scene = root.integer("scene", 0);
        const JVal* roots = jscenes->arr[(size_t)scene < jscenes->arr.size() ? (size_t)scene : 0].find("nodes");
        if (roots)
            for (size_t k = roots->arr.size(); k-- > 0;) stack.push_back((long)roots->arr[k].num);
        size_t visited = 0;
        while (!stack.empty()) {
            const long n = stack.back();
            stack.pop_back();
            if (n < 0 || (size_t)n >= jnodes->arr.size() || ++visited > 4 * jnodes->arr.size() + 16) return fail(std::string(path) + ": bad node graph");
            const JVal& node = jnodes->arr[n];
            const long m = node.integer("mesh", -1);
            if (m >= 0) {
                if ((size_t)m >= jmeshes->arr.size()) return fail(std::string(path) + ": node names a missing mesh");
                if (!add_mesh(jmeshes->arr[m])) return fail(std::string(path) + ": " + why);
            }
            if (const JVal* ch = node.find("children"))
                for (size_t k = ch->arr.size(); k-- > 0;) stack.push_back((long)ch->arr[k].num);
        }
    } else {
        for (const JVal& mesh : jmeshes->arr)
            if (!add_mesh(mesh)) return fail(std::string(path) + ": " + why);
    }
    if (M->indices.empty()) return fail(std::string("no triangles in ") + path);
    M->mesh_ids.resize(M->indices.size() / 3);
    for (size_t m = 0; m < M->mesh_first_index.size(); ++m) {
        const size_t lo = M->mesh_first_index[m] / 3, hi = (m + 1 < M->mesh_first_index.size() ? M->mesh_first_index[m + 1] : M->indices.size()) / 3;
        for (size_t t = lo; t < hi; ++t) M->mesh_ids[t] = first_mesh_number + (std::int32_t)m;
    }
    *out = owner.release();
    return CNDL_OK;
} catch (const std::bad_alloc&) {
    return CNDL_ERR_OOM;
} catch (...) {
    return CNDL_ERR_INVALID;
}

// By extension: .obj, .gltf, .glb (FileLoader::LoadModelFile, ModelFileLoader.cpp:229-238).
int cndl_model_load(const char* path, int32_t first_mesh_number, cndl_model** out, char* err, size_t err_cap) {
    if (!path) return CNDL_ERR_INVALID;
    const std::string p(path);
    auto ends = [&](const char* e) { const size_t n = std::strlen(e); return p.size() >= n && p.compare(p.size() - n, n, e) == 0; };
    if (ends(".gltf") || ends(".glb") || ends(".GLTF") || ends(".GLB")) return cndl_model_load_gltf(path, first_mesh_number, out, err, err_cap);
    return cndl_model_load_obj(path, first_mesh_number, out, err, err_cap);
}

void cndl_model_free(cndl_model* m) { delete m; }
size_t cndl_model_vertex_count(const cndl_model* m) { return m ? m->vertices.size() : 0; }
size_t cndl_model_index_count(const cndl_model* m) { return m ? m->indices.size() : 0; }
size_t cndl_model_mesh_count(const cndl_model* m) { return m ? m->mesh_first_index.size() : 0; }
const cndl_vertex* cndl_model_vertices(const cndl_model* m) { return m ? m->vertices.data() : nullptr; }
const uint32_t* cndl_model_indices(const cndl_model* m) { return m ? m->indices.data() : nullptr; }
const int32_t* cndl_model_mesh_ids(const cndl_model* m) { return m ? m->mesh_ids.data() : nullptr; }
const char* cndl_model_mesh_name(const cndl_model* m, size_t mesh) { return (m && mesh < m->mesh_names.size()) ? m->mesh_names[mesh].c_str() : ""; }
const char* cndl_model_mesh_albedo_path(const cndl_model* m, size_t mesh) { return (m && mesh < m->mesh_albedo.size()) ? m->mesh_albedo[mesh].c_str() : ""; }
const char* cndl_model_mesh_normal_path(const cndl_model* m, size_t mesh) { return (m && mesh < m->mesh_normal.size()) ? m->mesh_normal[mesh].c_str() : ""; }
int cndl_model_mesh_color(const cndl_model* m, size_t mesh, float rgb[3]) {
    if (!m || !rgb || mesh >= m->mesh_color.size()) return CNDL_ERR_INVALID;
    for (int k = 0; k < 3; ++k) rgb[k] = m->mesh_color[mesh][(size_t)k];
    return CNDL_OK;
}

// RayIntersector::GenerateMeshTextureReferences (Intersector.h:367-402): every handle not seen before takes the next index of the
// shader's Textures[] array — whether or not its path was found — and the table entry holds that index only for a valid path.
int cndl_generate_texture_references(const cndl_mesh_material* materials, size_t n, cndl_texture_reference* out, uint64_t* handles, size_t handles_cap,
                                     size_t* n_handles) try {
    if ((n && (!materials || !out)) || (handles_cap && !handles)) return CNDL_ERR_INVALID;
    std::map<std::uint64_t, int> index_of;  // m_TextureHandleReferenceMap
    int last = 0;
    for (size_t i = 0; i < n; ++i) {
        const cndl_mesh_material& mm = materials[i];
        if (index_of.find(mm.albedo_handle) == index_of.end()) index_of[mm.albedo_handle] = last++;
        if (index_of.find(mm.normal_handle) == index_of.end()) index_of[mm.normal_handle] = last++;
        cndl_texture_reference r;
        std::memset(&r, 0, sizeof(r));
        r.model_color[0] = mm.model_color[0]; r.model_color[1] = mm.model_color[1]; r.model_color[2] = mm.model_color[2]; r.model_color[3] = 1.0f;
        r.albedo = mm.albedo_valid ? index_of[mm.albedo_handle] : -1;
        r.normal = mm.normal_valid ? index_of[mm.normal_handle] : -1;
        out[i] = r;
    }
    for (const auto& kv : index_of)
        if ((size_t)kv.second < handles_cap) handles[kv.second] = kv.first;
    if (n_handles) *n_handles = (size_t)last;
    return CNDL_OK;
} catch (...) {
    return CNDL_ERR_OOM;
}

int cndl_add_model(cndl_ctx* ctx, uint32_t object_id, const cndl_model* m, const cndl_build_opts* opts) {
    if (!ctx || !m) return CNDL_ERR_INVALID;
    return cndl_add_object(ctx, object_id, m->vertices.data(), m->vertices.size(), m->indices.data(), m->indices.size(), m->mesh_ids.data(), opts);
}

}  // extern "C"
