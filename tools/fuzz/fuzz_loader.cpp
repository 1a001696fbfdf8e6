// Mutation fuzzer for the host-side model readers (candela_b200/csrc/model_loader.cu is plain host C++): built with
// -fsanitize=address,undefined by tools/fuzz/run_fuzz_loader.sh.  Seeds are small valid OBJ / glTF / GLB files written by
// tools/fuzz/make_seeds.py; every iteration mutates one seed (byte flips, truncation, splices, numeric tokens replaced by
// extreme values), writes it next to the seed's side files and calls cndl_model_load.  Any sanitizer report aborts the run.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "../../include/candela_b200.h"

// model_loader.cu references the scene calls through cndl_add_model only; the fuzzer never reaches them.
extern "C" int cndl_add_object(cndl_ctx*, uint32_t, const cndl_vertex*, size_t, const uint32_t*, size_t, const int32_t*, const cndl_build_opts*) { return CNDL_ERR_INVALID; }

static std::vector<unsigned char> slurp(const std::string& p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static const char* kTokens[] = {"-1", "0", "4294967295", "2147483647", "-2147483648", "18446744073709551615", "1e38", "-1e38", "nan", "inf", "1e-45",
                                "999999999999", "5126", "5123", "5121", "5125", "\"VEC3\"", "\"SCALAR\"", "\"VEC2\"", "\"MAT4\"", "[]", "{}", "null", "\"\"", "//", "1/2/3", "-1/-1/-1",
                                "1//1", "0/0/0"};

static void mutate(std::vector<unsigned char>& d, std::mt19937_64& rng, bool text) {
    const int n_mut = 1 + (int)(rng() % 4);
    for (int k = 0; k < n_mut && !d.empty(); ++k) {
        switch (rng() % (text ? 7 : 5)) {
            case 0: d[rng() % d.size()] ^= (unsigned char)(1u << (rng() % 8)); break;
            case 1: d[rng() % d.size()] = (unsigned char)rng(); break;
            case 2: d.resize(rng() % d.size()); break;                                       // truncate
            case 3: {                                                                          // splice a block over another place
                const size_t a = rng() % d.size(), b = rng() % d.size(), len = std::min<size_t>(rng() % 64, d.size() - std::max(a, b));
                std::memmove(d.data() + a, d.data() + b, len);
                break;
            }
            case 4: {                                                                          // 32-bit little-endian field -> extreme value
                if (d.size() < 4) break;
                const size_t a = rng() % (d.size() - 3);
                const uint32_t v[] = {0u, 1u, 0xFFFFFFFFu, 0x7FFFFFFFu, 0x80000000u, (uint32_t)d.size(), (uint32_t)d.size() + 1u, 12u};
                const uint32_t x = v[rng() % 8];
                std::memcpy(d.data() + a, &x, 4);
                break;
            }
            default: {                                                                         // replace a numeric / word token by an extreme one
                size_t a = rng() % d.size();
                while (a < d.size() && !(std::isdigit(d[a]) || d[a] == '-')) ++a;
                size_t b = a;
                while (b < d.size() && (std::isdigit(d[b]) || d[b] == '-' || d[b] == '.' || d[b] == 'e' || d[b] == '/')) ++b;
                const std::string t = kTokens[rng() % (sizeof(kTokens) / sizeof(kTokens[0]))];
                if (a < d.size()) {
                    d.erase(d.begin() + (long)a, d.begin() + (long)b);
                    d.insert(d.begin() + (long)a, t.begin(), t.end());
                }
                break;
            }
        }
    }
}

static int load_one(const std::string& path, long& ok, long& refused) {
    cndl_model* m = nullptr;
    char err[256] = {0};
    const int rc = cndl_model_load(path.c_str(), 0, &m, err, sizeof err);
    if (rc == CNDL_OK && m) {
        // touch everything a caller would read
        size_t sink = 0;
        const cndl_vertex* v = cndl_model_vertices(m);
        const uint32_t* idx = cndl_model_indices(m);
        const int32_t* ids = cndl_model_mesh_ids(m);
        const size_t nv = cndl_model_vertex_count(m), ni = cndl_model_index_count(m);
        for (size_t k = 0; k < ni; ++k) {
            if (idx[k] >= nv) { std::fprintf(stderr, "%s: index %u of %zu vertices\n", path.c_str(), idx[k], nv); return 5; }
            sink += idx[k];
        }
        for (size_t k = 0; k < ni / 3; ++k) sink += (size_t)ids[k];
        const unsigned char* vb = reinterpret_cast<const unsigned char*>(v);
        for (size_t k = 0; k < nv * sizeof(cndl_vertex); ++k) sink += vb[k];
        for (size_t k = 0; k < cndl_model_mesh_count(m); ++k) {
            float c[3];
            sink += std::strlen(cndl_model_mesh_name(m, k)) + std::strlen(cndl_model_mesh_albedo_path(m, k)) + std::strlen(cndl_model_mesh_normal_path(m, k));
            cndl_model_mesh_color(m, k, c);
        }
        volatile size_t keep = sink;
        (void)keep;
        ++ok;
        cndl_model_free(m);
    } else {
        ++refused;
        if (m) { std::fprintf(stderr, "error return with a model\n"); return 4; }
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 3 && std::string(argv[1]) == "files") {   // fuzz_loader files <path> ...: load crafted files as they are (tools/fuzz/craft_gltf.py)
        long ok = 0, refused = 0;
        for (int k = 2; k < argc; ++k)
            if (int rc = load_one(argv[k], ok, refused)) return rc;
        std::printf("fuzz_loader: %d crafted files, %ld loaded, %ld refused, no sanitizer report\n", argc - 2, ok, refused);
        return 0;
    }
    if (argc < 4) {
        std::fprintf(stderr, "usage: fuzz_loader <seed dir> <iterations> <rng seed>\n");
        return 2;
    }
    const std::string dir = argv[1];
    const long iters = std::atol(argv[2]);
    std::mt19937_64 rng((uint64_t)std::atoll(argv[3]));
    const char* seeds[] = {"tri.obj", "mtl.obj", "scene.gltf", "scene_uri.gltf", "scene.glb"};
    long ok = 0, refused = 0;
    for (long it = 0; it < iters; ++it) {
        const std::string name = seeds[rng() % 5];
        std::vector<unsigned char> d = slurp(dir + "/" + name);
        if (d.empty()) { std::fprintf(stderr, "missing seed %s\n", name.c_str()); return 2; }
        const bool text = name.find(".glb") == std::string::npos;
        if (it >= 5) mutate(d, rng, text);                                                  // the first rounds load the seeds unchanged
        const std::string ext = name.substr(name.rfind('.'));
        const std::string path = dir + "/fuzz_case" + ext;
        { std::ofstream o(path, std::ios::binary); o.write(reinterpret_cast<const char*>(d.data()), (std::streamsize)d.size()); }
        const long refused_before = refused;
        if (int rc = load_one(path, ok, refused)) return rc;
        if (it < 5 && refused != refused_before) { std::fprintf(stderr, "seed %s refused\n", name.c_str()); return 3; }
    }
    std::printf("fuzz_loader: %ld cases, %ld loaded, %ld refused, no sanitizer report\n", iters, ok, refused);
    return 0;
}
