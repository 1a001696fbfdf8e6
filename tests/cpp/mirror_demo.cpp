// Drives include/candela_b200/Intersector.hpp the way Pipeline.cpp:1019-1028,:1250-1251 drives the
// reference's RayIntersector: Initialize, AddObject, BufferData, PushEntities, BufferEntities, then a
// batch of rays.  Prints "OK <hits> <checksum>" on success; exits 3 when there is no GPU (the
// backend has no CPU fallback).  The Object/Mesh/Entity structs below stand in for the engine's.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "candela_b200/Intersector.hpp"

struct Mesh {
    std::vector<Candela::Vertex> m_Vertices;
    std::vector<unsigned> m_Indices;
    int GlobalMeshNumber = 0;
};
struct Object {
    std::vector<Mesh> m_Meshes;
    unsigned m_ObjectID = 2;  // Object.cpp:7-9: ids start at 2
    unsigned GetID() const { return m_ObjectID; }
};
struct Entity {
    const Object* m_Object;
    float m_Model[4][4];
    float m_EmissiveAmount = 0.0f, m_TranslucencyAmount = 0.0f;
};

// the demo links only against libcandela_b200.so: the device count comes from trying to open a context on device 1
static void cudaGetDeviceCountShim(int* n) {
    cndl_ctx* c = nullptr;
    *n = 1;
    if (cndl_create(&c, CNDL_STACKLESS, 1) == CNDL_OK) { *n = 2; cndl_destroy(c); }
}

int main() {
    Object obj;
    for (int m = 0; m < 2; ++m) {  // two meshes: a floor grid and a wall grid
        Mesh mesh;
        mesh.GlobalMeshNumber = 5 + m;
        const int n = 12;
        for (int i = 0; i <= n; ++i)
            for (int j = 0; j <= n; ++j) {
                Candela::Vertex v{};
                v.position[0] = (float)i;
                v.position[1] = m == 0 ? 0.0f : (float)j;
                v.position[2] = m == 0 ? (float)j : 0.0f;
                v.position[3] = 1.0f;
                mesh.m_Vertices.push_back(v);
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                const unsigned a = i * (n + 1) + j, b = a + n + 1;
                const unsigned idx[6] = {a, b, b + 1, a, b + 1, a + 1};
                mesh.m_Indices.insert(mesh.m_Indices.end(), idx, idx + 6);
            }
        obj.m_Meshes.push_back(mesh);
    }
    try {
        Candela::RayIntersector<Candela::BVH::StacklessTraversalNode> Intersector;
        try {
            Intersector.Initialize();
        } catch (const char* e) {
            std::printf("NO_GPU %s\n", e);
            return 3;
        }
        Intersector.AddObject(obj);
        Intersector.BufferData(true);
        Entity ent{&obj, {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}};
        std::vector<Entity*> list{&ent};
        Intersector.PushEntities(list);
        Intersector.BufferEntities();
        std::vector<Candela::Ray> rays;
        for (int i = 0; i < 100; ++i) {
            Candela::Ray r{{0.5f + 0.11f * i, 3.0f, 0.5f + 0.07f * i}, 0.0f, {0.0f, -1.0f, 0.0f}, 0.0f};
            rays.push_back(r);
        }
        std::vector<Candela::RayHit> hits(rays.size());
        Intersector.IntersectRays(rays.data(), rays.size(), hits.data());
        int n_hit = 0;
        double sum = 0;
        for (auto& h : hits)
            if (h.T > 0) { ++n_hit; sum += h.T; if (h.Mesh != 5) { std::printf("BAD mesh %d\n", h.Mesh); return 1; } }
        Intersector.FetchCPUData();
        std::printf("OK %d %.3f nodes=%zu tris=%zu\n", n_hit, sum, Intersector.m_BVHNodes.size(), Intersector.m_BVHTriangles.size());
        // GetData on the hit records and the collide query (Physics::CollideBox / CollidePoint)
        std::vector<Candela::HitData> data(hits.size());
        Intersector.GetData(hits.data(), hits.size(), data.data());
        int n_data = 0;
        for (std::size_t i = 0; i < hits.size(); ++i)
            if (hits[i].T > 0 && data[i].Mesh == 5) ++n_data;
        const float on_floor[3] = {1.0f, 0.0f, 1.0f}, in_air[3] = {1.0f, 1.5f, 1.0f};
        const float bmin[3] = {0.5f, -0.2f, 0.5f}, bmax[3] = {0.8f, 0.2f, 0.8f};
        std::printf("EXTRA data=%d collide=%d%d%d\n", n_data, (int)Candela::Physics::CollidePoint(on_floor, Intersector),
                    (int)Candela::Physics::CollidePoint(in_air, Intersector), (int)Candela::Physics::CollideBox(bmin, bmax, Intersector));
        // GenerateMeshTextureReferences + GetData with the Albedo decision: mesh 5 textured (valid handle), the others plain
        {
            std::vector<Candela::FileLoader::_MeshMaterialData> mats(6);
            for (int m = 0; m < 6; ++m) mats[m] = {0xA0u + m, 0xB0u, m == 5 ? 1 : 0, 0, {0.1f * m, 0.5f, 1.0f}, 0.0f};
            Intersector.GenerateMeshTextureReferences(mats);
            std::vector<Candela::HitMaterial> mat(hits.size());
            Intersector.GetData(hits.data(), hits.size(), mat.data());
            int n_tex = 0, n_same = 0;
            for (std::size_t i = 0; i < hits.size(); ++i) {
                if (hits[i].T > 0 && mat[i].AlbedoRef == Intersector.m_MeshTextureReferences[5].Albedo && mat[i].Albedo[0] == 0.0f) ++n_tex;
                if (std::memcmp(&mat[i], &data[i], sizeof(data[i])) == 0) ++n_same;
            }
            std::printf("MATERIAL tex=%d same=%d handles=%zu ref5=%d ref0=%d\n", n_tex, n_same, Intersector.m_TextureHandles.size(),
                        Intersector.m_MeshTextureReferences[5].Albedo, Intersector.m_MeshTextureReferences[0].Albedo);
        }
        // BVH::BuildBVH as a free function (BVHConstructor.h:86-87): same bytes as the intersector's own buffers for its first object
        {
            std::vector<Candela::BVH::FlattenedNode> Nodes;
            std::vector<Candela::Vertex> Vertices;
            std::vector<Candela::BVH::Triangle> Triangles;
            Candela::BVH::Node* Root = Candela::BVH::BuildBVH(obj, Nodes, Vertices, Triangles, 0);
            const bool same = Root == nullptr && Nodes.size() == Intersector.m_BVHNodes.size() && Triangles.size() == Intersector.m_BVHTriangles.size() &&
                              Vertices.size() == Intersector.m_BVHVertices.size() &&
                              std::memcmp(Nodes.data(), Intersector.m_BVHNodes.data(), Nodes.size() * sizeof(Nodes[0])) == 0 &&
                              std::memcmp(Triangles.data(), Intersector.m_BVHTriangles.data(), Triangles.size() * sizeof(Triangles[0])) == 0;
            std::vector<Candela::BVH::FlattenedStackNode> StackNodes;
            std::vector<Candela::Vertex> V2;
            std::vector<Candela::BVH::Triangle> T2;
            Candela::BVH::BuildBVH(obj, StackNodes, V2, T2, 1000);
            std::printf("BUILDBVH same=%d stack_nodes=%zu tris=%zu\n", (int)same, StackNodes.size(), T2.size());
        }
        // a diffuse frame (IntersectDiffuse: camera rays -> hits -> cosine-hemisphere rays -> hits, all on the device), and the same
        // frame from a second intersector that drives two contexts (two GPUs when the box has them, else the same GPU twice):
        // tiles dealt round-robin, records gathered into the first device's frame — identical bytes
        {
            const int W = 160, H = 96;
            // camera at (6, 5, 6) looking along -y/-z onto the floor and the wall: inverse view = rotation * translation, column-major
            const float InverseView[16] = {1, 0, 0, 0, 0, 0.70710678f, -0.70710678f, 0, 0, 0.70710678f, 0.70710678f, 0, 6, 5, 6, 1};
            const float InverseProjection[16] = {1.6666666f, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, -24.99f, 0, 0, -1, 25.01f};
            std::vector<cndl_hit16> a((size_t)W * H), b((size_t)W * H);
            Intersector.IntersectDiffuse(a.data(), W, H, InverseView, InverseProjection, 42u);
            int n_d = 0;
            for (auto& h : a) n_d += h.t > 0 ? 1 : 0;
            int n_dev = 0;
            cudaGetDeviceCountShim(&n_dev);
            Candela::RayIntersector<Candela::BVH::StacklessTraversalNode> Multi;
            Multi.Initialize(std::vector<int>{0, n_dev > 1 ? 1 : 0});
            Multi.AddObject(obj);
            Multi.BufferData(true);
            Multi.PushEntities(list);
            Multi.BufferEntities();
            Multi.IntersectDiffuse(b.data(), W, H, InverseView, InverseProjection, 42u);
            std::printf("FRAME hits=%d devices=%d same=%d\n", n_d, Multi.DeviceCount(), (int)(std::memcmp(a.data(), b.data(), a.size() * sizeof(a[0])) == 0));
        }
        // pushing an entity of an unknown object must throw the reference's message
        Object other;
        other.m_ObjectID = 77;
        Entity bad{&other, {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}};
        try {
            Intersector.PushEntity(bad);
            std::printf("BAD no throw\n");
            return 1;
        } catch (const char* e) {
            std::printf("THROW %s\n", e);
        }
    } catch (const char* e) {
        std::printf("ERROR %s\n", e);
        return 1;
    }
    return 0;
}
