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
               *jnodes = root.find("nodes"), *jscenes = root.find("scenes");
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
            Accessor pos, nor, uv, idx;
            if (!accessor(attrs->integer("POSITION", -1), pos, why) || pos.width != 3) { if (why.empty()) why = "POSITION must be VEC3"; return false; }
            const bool has_n = attrs->find("NORMAL") != nullptr, has_uv = attrs->find("TEXCOORD_0") != nullptr, has_i = prim.find("indices") != nullptr;
            if (has_n && (!accessor(attrs->integer("NORMAL", -1), nor, why) || nor.count < pos.count)) { if (why.empty()) why = "NORMAL shorter than POSITION"; return false; }
            if (has_uv && (!accessor(attrs->integer("TEXCOORD_0", -1), uv, why) || uv.count < pos.count)) { if (why.empty()) why = "TEXCOORD_0 shorter than POSITION"; return false; }
            if (has_i && !accessor(prim.integer("indices", -1), idx, why)) return false;
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
            if (has_n) {  // indexed as in the file
                for (std::uint32_t v = 0; v < pos.count; ++v) {
                    const float n[3] = {nor.get(v, 0), nor.get(v, 1), nor.get(v, 2)};
                    M->vertices.push_back(make_vertex(v, n));
                }
                for (std::uint32_t c : corners) M->indices.push_back(first_vertex + c);
            } else {  // aiProcess_GenNormals: flat shading, one vertex per corner
                for (size_t k = 0; k + 2 < corners.size() + 0 && k + 2 < corners.size(); k += 3) {
                    float p[3][3];
                    for (int c = 0; c < 3; ++c)
                        for (int d = 0; d < 3; ++d) p[c][d] = pos.get(corners[k + c], d);
                    const float e1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]}, e2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
                    float fn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                    const float len = std::sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
                    if (len > 0.0f) { fn[0] /= len; fn[1] /= len; fn[2] /= len; }
                    for (int c = 0; c < 3; ++c) {
                        M->indices.push_back((std::uint32_t)M->vertices.size());
                        M->vertices.push_back(make_vertex(corners[k + c], fn));
                    }
                }
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

int cndl_add_model(cndl_ctx* ctx, uint32_t object_id, const cndl_model* m, const cndl_build_opts* opts) {
    if (!ctx || !m) return CNDL_ERR_INVALID;
    return cndl_add_object(ctx, object_id, m->vertices.data(), m->vertices.size(), m->indices.data(), m->indices.size(), m->mesh_ids.data(), opts);
}

}  // extern "C"
