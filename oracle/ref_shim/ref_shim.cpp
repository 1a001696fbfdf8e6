// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// C-ABI shim around the UNMODIFIED reference builder (Candela::BVH::BuildBVH,
// /root/reference/Source/Core/BVH/BVHConstructor.cpp:951 and :1032).  The
// reference sources are compiled from where they lie under /root/reference by
// oracle/Makefile (target `ref`); nothing from them is copied into this repo.
// The resulting oracle/_ref/libcandela_ref.so is used by tests/ and by
// tests/golden/make_golden.py to pin the oracle's builder restatement
// byte-for-byte, and (optionally) by bench.py's cpu_baseline leg.
//
// The reference's Mesh constructor and GL buffer destructors call OpenGL
// through glad's function pointers (Mesh.cpp:5-15); there is no GL context
// here, so those pointers are aimed at no-op stubs before any Object is made.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#include <glad/glad.h>

#include "BVH/BVHConstructor.h"
#include "Object.h"
#include "Physics.h"
#include <glm/glm.hpp>
#include <glm/gtc/packing.hpp>

#include <new>

namespace {

void APIENTRY stub_gen(GLsizei n, GLuint* ids) { for (GLsizei i = 0; i < n; ++i) ids[i] = 1; }
void APIENTRY stub_del(GLsizei, const GLuint*) {}
void APIENTRY stub_bind_buffer(GLenum, GLuint) {}
void APIENTRY stub_bind_vao(GLuint) {}
void APIENTRY stub_attrib(GLuint, GLint, GLenum, GLboolean, GLsizei, const void*) {}
void APIENTRY stub_attrib_i(GLuint, GLint, GLenum, GLsizei, const void*) {}
void APIENTRY stub_enable(GLuint) {}

void install_gl_stubs() {
    glad_glGenBuffers = stub_gen;
    glad_glGenVertexArrays = stub_gen;
    glad_glDeleteBuffers = stub_del;
    glad_glDeleteVertexArrays = stub_del;
    glad_glDeleteTextures = stub_del;
    glad_glBindBuffer = stub_bind_buffer;
    glad_glBindVertexArray = stub_bind_vao;
    glad_glVertexAttribPointer = stub_attrib;
    glad_glVertexAttribIPointer = stub_attrib_i;
    glad_glEnableVertexAttribArray = stub_enable;
}

std::vector<Candela::BVH::FlattenedNode> g_nodes_stackless;
std::vector<Candela::BVH::FlattenedStackNode> g_nodes_stack;
std::vector<Candela::Vertex> g_verts;
std::vector<Candela::BVH::Triangle> g_tris;
int g_format = -1;

}  // namespace

extern "C" {

// Sizes of the reference's own structs, so the tests can check the layout
// contract (32/16/32/64) against the real thing.
void ref_sizes(int out[4]) {
    out[0] = (int)sizeof(Candela::Vertex);
    out[1] = (int)sizeof(Candela::BVH::Triangle);
    out[2] = (int)sizeof(Candela::BVH::FlattenedNode);
    out[3] = (int)sizeof(Candela::BVH::FlattenedStackNode);
}

// Builds one object made of `n_meshes` meshes with the reference builder.
//   format 0 = FlattenedNode (stackless; the reference flips children at
//   random from std::random_device), 1 = FlattenedStackNode.
//   verts: concatenated 32-byte Vertex records, mesh after mesh.
//   indices: concatenated mesh-local indices.
// Returns 0, or -1 when the reference would crash (fewer than 100 triangles:
// `TotalIterations % StatusFrequency` divides by zero, BVHConstructor.cpp:387,:449).
int ref_build(int format, int n_meshes, const void* verts, const uint64_t* mesh_vertex_counts,
              const uint32_t* indices, const uint64_t* mesh_index_counts,
              const int32_t* mesh_global_numbers, int t_offset,
              uint64_t* out_n_nodes, uint64_t* out_n_tris, uint64_t* out_n_verts) {
    install_gl_stubs();
    uint64_t total_idx = 0;
    for (int m = 0; m < n_meshes; ++m) total_idx += mesh_index_counts[m];
    if (total_idx < 300) return -1;

    Candela::Object object;
    const Candela::Vertex* vp = static_cast<const Candela::Vertex*>(verts);
    const uint32_t* ip = indices;
    for (int m = 0; m < n_meshes; ++m) {
        Candela::Mesh& mesh = object.GenerateMesh();
        mesh.m_Vertices.assign(vp, vp + mesh_vertex_counts[m]);
        mesh.m_Indices.assign(ip, ip + mesh_index_counts[m]);
        mesh.GlobalMeshNumber = mesh_global_numbers[m];
        vp += mesh_vertex_counts[m];
        ip += mesh_index_counts[m];
    }

    g_nodes_stackless.clear();
    g_nodes_stack.clear();
    g_verts.clear();
    g_tris.clear();
    g_format = format;

    // The reference prints progress to std::cout; silence it for the call.
    std::ostringstream sink;
    std::streambuf* saved = std::cout.rdbuf(sink.rdbuf());
    if (format == 0) {
        Candela::BVH::BuildBVH(object, g_nodes_stackless, g_verts, g_tris, t_offset);
    } else {
        Candela::BVH::BuildBVH(object, g_nodes_stack, g_verts, g_tris, t_offset);
    }
    std::cout.rdbuf(saved);

    *out_n_nodes = format == 0 ? g_nodes_stackless.size() : g_nodes_stack.size();
    *out_n_tris = g_tris.size();
    *out_n_verts = g_verts.size();
    return 0;
}

// Copies the buffers of the last ref_build out. Any pointer may be null.
void ref_fetch(void* nodes, void* tris, void* verts) {
    if (nodes) {
        if (g_format == 0)
            std::memcpy(nodes, g_nodes_stackless.data(), g_nodes_stackless.size() * sizeof(g_nodes_stackless[0]));
        else
            std::memcpy(nodes, g_nodes_stack.data(), g_nodes_stack.size() * sizeof(g_nodes_stack[0]));
    }
    if (tris) std::memcpy(tris, g_tris.data(), g_tris.size() * sizeof(g_tris[0]));
    if (verts) std::memcpy(verts, g_verts.data(), g_verts.size() * sizeof(g_verts[0]));
}

// Physics::CollideBox (Physics.cpp:203-228) of the UNMODIFIED reference on caller-supplied stackless buffers.  The
// function only reads the intersector's public vectors (Intersector.h:96-99); an intersector cannot be constructed
// without a GL context (its members compile shaders), so the four vectors are placement-constructed inside raw
// storage of the right type and nothing else of the object is touched.
int ref_collide_box(const void* nodes, uint64_t n_nodes, const void* tris, uint64_t n_tris, const void* verts, uint64_t n_verts,
                    const void* entities, uint64_t n_entities, const float* boxes, uint64_t n, int32_t* out) {
    using namespace Candela;
    typedef RayIntersector<BVH::StacklessTraversalNode> RI;
    alignas(64) static unsigned char storage[sizeof(RI)];
    RI* ri = reinterpret_cast<RI*>(storage);
    auto* np_ = static_cast<const BVH::StacklessTraversalNode*>(nodes);
    auto* tp = static_cast<const BVH::Triangle*>(tris);
    auto* vp = static_cast<const Vertex*>(verts);
    auto* ep = static_cast<const BVHEntity*>(entities);
    new (&ri->m_BVHNodes) std::vector<BVH::StacklessTraversalNode>(np_, np_ + n_nodes);
    new (&ri->m_BVHTriangles) std::vector<BVH::Triangle>(tp, tp + n_tris);
    new (&ri->m_BVHVertices) std::vector<Vertex>(vp, vp + n_verts);
    new (&ri->m_BVHEntities) std::vector<BVHEntity>(ep, ep + n_entities);
    for (uint64_t q = 0; q < n; ++q) {
        const glm::vec3 mn(boxes[8 * q], boxes[8 * q + 1], boxes[8 * q + 2]), mx(boxes[8 * q + 4], boxes[8 * q + 5], boxes[8 * q + 6]);
        out[q] = Physics::CollideBox(mn, mx, *ri) ? 1 : 0;
    }
    using VN = std::vector<BVH::StacklessTraversalNode>; using VT = std::vector<BVH::Triangle>; using VV = std::vector<Vertex>; using VE = std::vector<BVHEntity>;
    ri->m_BVHNodes.~VN(); ri->m_BVHTriangles.~VT(); ri->m_BVHVertices.~VV(); ri->m_BVHEntities.~VE();
    return 0;
}

// glm::packHalf2x16 of the reference's vendored glm 0.9.8.5 (what ModelFileLoader.cpp:133-155 packs vertices with).
uint32_t ref_pack_half2x16(float x, float y) { return glm::packHalf2x16(glm::vec2(x, y)); }

}  // extern "C"
