// Scene ingest without Assimp: a Wavefront OBJ reader that produces what
// ModelFileLoader.cpp:101-185 hands to the intersector — 32-byte Vertex records with the normal, tangent
// and UV packed as half floats exactly like glm::packHalf2x16 (ModelFileLoader.cpp:133-155), object-local
// indices with the per-mesh vertex offset already applied (BVHConstructor.cpp:981-1002) and one
// GlobalMeshNumber per triangle (ModelFileLoader.cpp:104-105: a running counter, one per mesh).
// Host code only.  The reference imports with aiProcess_JoinIdenticalVertices | Triangulate | CalcTangentSpace |
// GenUVCoords | FlipUVs | GenNormals (ModelFileLoader.cpp:243-252).  This reader does the same where the
// intersector can see it: one mesh per `usemtl` / `o` / `g` run (Assimp: one mesh per material), polygons
// fan-triangulated, v flipped (uv.y = 1 - uv.y), a flat face normal generated for corners that name no `vn`, and
// vertices joined per mesh when their position index, UV index and normal agree.  Tangents are left zero
// (CalcTangentSpace feeds the raster pass only; GetData reads normal and UV).  The ORDER of the joined
// vertices is not Assimp's — which cannot matter: the builder and the traversal only ever see the arrays
// this loader returns.
#include <cerrno>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

}  // namespace

struct cndl_model {
    std::vector<cndl_vertex> vertices;
    std::vector<std::uint32_t> indices;   // object-local: mesh-local index + vertices of earlier meshes
    std::vector<std::int32_t> mesh_ids;   // one per triangle
    std::vector<std::uint32_t> mesh_first_vertex, mesh_first_index;
    std::vector<std::string> mesh_names;
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

    std::vector<float> P, N, T;  // v (3), vn (3), vt (2)
    // n >= 0: index into the file's `vn` list; n == -1: generated face normal, whose value is in g[]
    struct Key {
        long v, t, n;
        float g[3];
        bool operator==(const Key& o) const { return v == o.v && t == o.t && n == o.n && std::memcmp(g, o.g, sizeof(g)) == 0; }
    };
    struct KeyHash {
        size_t operator()(const Key& k) const {
            std::uint32_t b[3];
            std::memcpy(b, k.g, sizeof(b));
            return (size_t)k.v * 73856093u ^ (size_t)(k.t + 1) * 19349663u ^ (size_t)(k.n + 1) * 83492791u ^ b[0] ^ (b[1] * 31u) ^ (b[2] * 131u);
        }
    };
    std::unordered_map<Key, std::uint32_t, KeyHash> seen;  // per mesh
    bool mesh_open = false;
    std::string pending_name = "default";
    auto open_mesh = [&]() {
        M->mesh_first_vertex.push_back((std::uint32_t)M->vertices.size());
        M->mesh_first_index.push_back((std::uint32_t)M->indices.size());
        M->mesh_names.push_back(pending_name);
        seen.clear();
        mesh_open = true;
    };
    auto vertex_of = [&](Key k) -> std::uint32_t {
        auto it = seen.find(k);
        if (it != seen.end()) return it->second;
        cndl_vertex v;
        std::memset(&v, 0, sizeof(v));
        v.position[0] = P[3 * k.v]; v.position[1] = P[3 * k.v + 1]; v.position[2] = P[3 * k.v + 2]; v.position[3] = 1.0f;
        float nx = k.g[0], ny = k.g[1], nz = k.g[2], tu = 0.0f, tv = 0.0f;
        if (k.n >= 0) { nx = N[3 * k.n]; ny = N[3 * k.n + 1]; nz = N[3 * k.n + 2]; }
        if (k.t >= 0) { tu = T[2 * k.t]; tv = 1.0f - T[2 * k.t + 1]; }  // aiProcess_FlipUVs
        v.normal_tangent[0] = cndl_pack_half2x16(nx, ny);    // data.x = packHalf2x16(vnormal.xy)   (ModelFileLoader.cpp:150)
        v.normal_tangent[1] = cndl_pack_half2x16(nz, 0.0f);  // data.y = packHalf2x16(vnormal.z, vtan.x); no tangents in an OBJ
        v.normal_tangent[2] = cndl_pack_half2x16(0.0f, 0.0f);
        v.texcoords = cndl_pack_half2x16(tu, tv);            // :133-136, (0,0) when the mesh has no UVs (:148)
        const std::uint32_t idx = (std::uint32_t)M->vertices.size();  // object-local = mesh-local + vertices of earlier meshes
        M->vertices.push_back(v);
        seen.emplace(k, idx);
        return idx;
    };

    std::vector<char> line(1 << 16);
    long lineno = 0;
    std::string problem;
    while (std::fgets(line.data(), (int)line.size(), f)) {
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
            if (!mesh_open) open_mesh();
            std::vector<Key> corners;
            char* e = s + 1;
            while (true) {
                while (*e == ' ' || *e == '\t') ++e;
                if (*e == '\0' || *e == '\n' || *e == '\r' || *e == '#') break;
                Key k{0, -1, -1, {0.0f, 0.0f, 0.0f}};
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
            if (corners.size() >= 3) {  // aiProcess_GenNormals: corners without a `vn` get the face's normal
                const float* a = &P[3 * corners[0].v];
                const float* b = &P[3 * corners[1].v];
                const float* c = &P[3 * corners[2].v];
                const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
                float fn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                const float len = std::sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
                if (len > 0.0f) { fn[0] /= len; fn[1] /= len; fn[2] /= len; }
                for (auto& k : corners)
                    if (k.n < 0) { k.g[0] = fn[0]; k.g[1] = fn[1]; k.g[2] = fn[2]; }
            }
            for (size_t c = 1; c + 1 < corners.size(); ++c) {  // triangle fan, like aiProcess_Triangulate on convex polygons
                M->indices.push_back(vertex_of(corners[0]));
                M->indices.push_back(vertex_of(corners[c]));
                M->indices.push_back(vertex_of(corners[c + 1]));
            }
        } else if (std::strncmp(s, "usemtl", 6) == 0 || ((s[0] == 'o' || s[0] == 'g') && (s[1] == ' ' || s[1] == '\t'))) {
            char* e = s + (s[0] == 'u' ? 6 : 1);
            while (*e == ' ' || *e == '\t') ++e;
            std::string name(e);
            while (!name.empty() && (name.back() == '\n' || name.back() == '\r' || name.back() == ' ')) name.pop_back();
            pending_name = name.empty() ? "default" : name;
            // a new mesh starts at the next face; a mesh that has no face yet is simply renamed
            if (mesh_open && M->indices.size() > M->mesh_first_index.back()) mesh_open = false;
            else if (mesh_open) M->mesh_names.back() = pending_name;
        }
    }
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

void cndl_model_free(cndl_model* m) { delete m; }
size_t cndl_model_vertex_count(const cndl_model* m) { return m ? m->vertices.size() : 0; }
size_t cndl_model_index_count(const cndl_model* m) { return m ? m->indices.size() : 0; }
size_t cndl_model_mesh_count(const cndl_model* m) { return m ? m->mesh_first_index.size() : 0; }
const cndl_vertex* cndl_model_vertices(const cndl_model* m) { return m ? m->vertices.data() : nullptr; }
const uint32_t* cndl_model_indices(const cndl_model* m) { return m ? m->indices.data() : nullptr; }
const int32_t* cndl_model_mesh_ids(const cndl_model* m) { return m ? m->mesh_ids.data() : nullptr; }
const char* cndl_model_mesh_name(const cndl_model* m, size_t mesh) { return (m && mesh < m->mesh_names.size()) ? m->mesh_names[mesh].c_str() : ""; }

int cndl_add_model(cndl_ctx* ctx, uint32_t object_id, const cndl_model* m, const cndl_build_opts* opts) {
    if (!ctx || !m) return CNDL_ERR_INVALID;
    return cndl_add_object(ctx, object_id, m->vertices.data(), m->vertices.size(), m->indices.data(), m->indices.size(), m->mesh_ids.data(), opts);
}

}  // extern "C"
