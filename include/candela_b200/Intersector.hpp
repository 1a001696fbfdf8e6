// Header-only C++17 mirror of Candela's host-side intersector API over the C ABI of
// libcandela_b200.so.  It has the reference's class, method and public-member names
// (Source/Core/BVH/Intersector.h:60-124, BVHConstructor.h:61-87) so that engine code such as
//
//     Candela::RayIntersector<Candela::BVH::StacklessTraversalNode> Intersector;
//     Intersector.Initialize();
//     Intersector.AddObject(MainModel);
//     Intersector.BufferData(true);
//     ... per frame: Intersector.PushEntities(EntityRenderList); Intersector.BufferEntities();
//
// (Pipeline.cpp:1019-1028, :1250-1251) compiles against it unchanged.  The OpenGL-specific members
// (SSBO ids, BindEverything(shader), texture tables) have no meaning on a CUDA backend; their
// replacement is DeviceBuffers() (device pointers for the caller's own kernels) and the batch
// queries IntersectRays / IntersectRaysAny, which return the hit records the GLSL callers consumed.
//
// Object / Mesh / Entity are taken as template parameters ("duck typing"): anything with the
// reference's member names works, including the engine's own classes
// (Object::m_Meshes, Object::GetID(), Mesh::m_Vertices / m_Indices / GlobalMeshNumber,
//  Entity::m_Object, m_Model, m_EmissiveAmount, m_TranslucencyAmount).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../candela_b200.h"

namespace Candela {

struct Vertex {  // Utils/Vertex.h:7-12
    float position[4];
    std::uint32_t normal_tangent_data[3];
    std::uint32_t texcoords;
};

namespace BVH {
struct FBounds { float Min[4]; float Max[4]; };
struct FlattenedNode { float Min[4]; float Max[4]; };            // BVHConstructor.h:66-70
struct FlattenedStackNode { FBounds LBounds; FBounds RBounds; };  // BVHConstructor.h:72-77
struct Triangle { int PackedData[4]; };                           // BVHConstructor.h:79-84
typedef FlattenedNode StacklessTraversalNode;                     // Intersector.h:39
typedef FlattenedStackNode StackTraversalNode;                    // Intersector.h:40
struct Node;  // the reference returns its heap root (BVHConstructor.h:86-87), which every caller drops; here it is always null
}  // namespace BVH

namespace detail {
// The mesh concatenation of BuildBVH (BVHConstructor.cpp:981-1002): indices + running vertex offset, vertices appended, one
// GlobalMeshNumber per triangle.
template <typename ObjectT>
inline void ConcatenateMeshes(const ObjectT& object, std::vector<Vertex>& Vertices, std::vector<std::uint32_t>& MeshIndices,
                              std::vector<std::int32_t>& MeshReferences) {
    std::uint32_t IndexOffset = 0;
    for (const auto& Mesh : object.m_Meshes) {
        for (std::size_t x = 0; x < Mesh.m_Indices.size(); ++x) {
            MeshIndices.push_back(static_cast<std::uint32_t>(Mesh.m_Indices[x]) + IndexOffset);
            if (x % 3 == 0) MeshReferences.push_back(Mesh.GlobalMeshNumber);
        }
        const std::size_t at = Vertices.size();
        Vertices.resize(at + Mesh.m_Vertices.size());
        static_assert(sizeof(Mesh.m_Vertices[0]) == sizeof(Vertex), "Vertex must be the 32-byte record");
        if (!Mesh.m_Vertices.empty()) std::memcpy(&Vertices[at], &Mesh.m_Vertices[0], Mesh.m_Vertices.size() * sizeof(Vertex));
        IndexOffset += static_cast<std::uint32_t>(Mesh.m_Vertices.size());
    }
}
template <typename NodeT, typename ObjectT>
inline BVH::Node* BuildBVH(int Format, const ObjectT& object, std::vector<NodeT>& FlattenedNodes, std::vector<Vertex>& MeshVertices,
                           std::vector<BVH::Triangle>& FlattenedTris, int t_offset, int Device, const cndl_build_opts* Options) {
    std::vector<std::uint32_t> MeshIndices;
    std::vector<std::int32_t> MeshReferences;
    const std::size_t FirstVertex = MeshVertices.size();
    ConcatenateMeshes(object, MeshVertices, MeshIndices, MeshReferences);
    const std::size_t T = MeshIndices.size() / 3;
    if (T == 0) throw std::string("candela_b200: BuildBVH on an object without triangles");
    FlattenedNodes.resize(2 * T - 1);
    FlattenedTris.resize(T);
    std::size_t N = 0;
    const int rc = cndl_build_bvh(Format, Device, reinterpret_cast<const cndl_vertex*>(MeshVertices.data() + FirstVertex),
                                  MeshVertices.size() - FirstVertex, MeshIndices.data(), MeshIndices.size(), MeshReferences.data(), t_offset,
                                  Options, FlattenedNodes.data(), FlattenedNodes.size(), &N, reinterpret_cast<cndl_triangle*>(FlattenedTris.data()),
                                  nullptr);
    if (rc != CNDL_OK) throw std::string("candela_b200: cndl_build_bvh failed with status ") + std::to_string(rc);
    FlattenedNodes.resize(N);
    return nullptr;
}
}  // namespace detail

namespace BVH {
// BVH::BuildBVH (BVHConstructor.h:86-87; .cpp:951-1028 stackless, :1032-1108 stack), built on the GPU, buffers byte-identical:
// fills FlattenedNodes (leaf packs include t_offset) and FlattenedTris (object-local vertex indices), appends the object's
// vertices to MeshVertices.
template <typename ObjectT>
inline Node* BuildBVH(const ObjectT& object, std::vector<FlattenedNode>& FlattenedNodes, std::vector<Vertex>& MeshVertices,
                      std::vector<Triangle>& FlattenedTris, int t_offset, int Device = 0, const cndl_build_opts* Options = nullptr) {
    return detail::BuildBVH(CNDL_STACKLESS, object, FlattenedNodes, MeshVertices, FlattenedTris, t_offset, Device, Options);
}
template <typename ObjectT>
inline Node* BuildBVH(const ObjectT& object, std::vector<FlattenedStackNode>& FlattenedNodes, std::vector<Vertex>& MeshVertices,
                      std::vector<Triangle>& FlattenedTris, int t_offset, int Device = 0, const cndl_build_opts* Options = nullptr) {
    return detail::BuildBVH(CNDL_STACK, object, FlattenedNodes, MeshVertices, FlattenedTris, t_offset, Device, Options);
}
}  // namespace BVH

struct BVHEntity {  // Intersector.h:43-49
    float ModelMatrix[16];
    float InverseMatrix[16];
    int NodeOffset;
    int NodeCount;
    int Data[14];
};

struct RayHit { float T, U, V, W; int Mesh, TriangleIdx, Entity, Iters; };  // cndl_hit
struct HitData { float Normal[3]; float UV[2]; float Emissivity; float Alpha; int Mesh; };  // cndl_hit_attr: the outputs of GetData
struct Ray { float Origin[3]; float TMin; float Direction[3]; float TMax; }; // cndl_ray
struct HitMaterial { float Normal[3]; float UV[2]; float Emissivity; float Alpha; int Mesh; float Albedo[3]; int AlbedoRef; };  // cndl_hit_material
namespace BVH {
struct TextureReferences { float ModelColor[4]; int Albedo; int Normal; int Pad[2]; };  // Intersector.h:32-37
}
namespace FileLoader {
// ModelFileLoader.h:21-25 with the two paths already resolved by the texture cache (Texture.cpp:168-188): a 64-bit handle + found flag.
struct _MeshMaterialData { std::uint64_t AlbedoHandle, NormalHandle; int AlbedoValid, NormalValid; float ModelColor[3]; float Pad; };  // cndl_mesh_material
}

static_assert(sizeof(Vertex) == 32 && sizeof(BVH::Triangle) == 16 && sizeof(BVH::FlattenedNode) == 32 &&
                  sizeof(BVH::FlattenedStackNode) == 64 && sizeof(BVHEntity) == 192 && sizeof(RayHit) == 32 && sizeof(Ray) == 32 &&
                  sizeof(HitData) == 32 && sizeof(HitMaterial) == 48 && sizeof(BVH::TextureReferences) == 32 && sizeof(FileLoader::_MeshMaterialData) == 40,
              "record layouts are the contract (SURVEY.md §8a)");

template <typename T>
class RayIntersector {
public:
    RayIntersector() {
        // Intersector.h:138-148
        if (std::is_same<T, BVH::FlattenedStackNode>::value) m_Stackless = false;
        else if (std::is_same<T, BVH::FlattenedNode>::value) m_Stackless = true;
        else throw "\nTemplate <T> Passed to RayIntersector can only be of type BVH::FlattenedStackNode or BVH::FlattenedNode>!";
    }
    ~RayIntersector() {
        if (m_Multi) cndl_multi_destroy(m_Multi);  // owns its per-device contexts, m_Ctx among them
        else if (m_Ctx) cndl_destroy(m_Ctx);
    }
    RayIntersector(const RayIntersector&) = delete;
    RayIntersector& operator=(const RayIntersector&) = delete;

    // Intersector.h:154 compiles the trace shader; here it opens the CUDA context on `Device`.
    void Initialize(int Device = 0) {
        if (m_Ctx) return;
        if (cndl_create(&m_Ctx, m_Stackless ? CNDL_STACKLESS : CNDL_STACK, Device) != CNDL_OK)
            throw "candela_b200: no usable sm_100 CUDA device (there is no CPU fallback)";
    }

    // Several GPUs behind the same object: the BVH is built once on Devices[0] and replicated device to device,
    // TraceFrame deals the screen tiles round-robin and gathers the records into Devices[0]'s frame over NVLink peer memory.
    // Every other query runs on Devices[0].
    void Initialize(const std::vector<int>& Devices) {
        if (m_Ctx) return;
        if (Devices.size() <= 1) { Initialize(Devices.empty() ? 0 : Devices[0]); return; }
        if (cndl_multi_create(&m_Multi, m_Stackless ? CNDL_STACKLESS : CNDL_STACK, Devices.data(), (int)Devices.size()) != CNDL_OK)
            throw "candela_b200: no usable sm_100 CUDA device (there is no CPU fallback)";
        m_Ctx = cndl_multi_context(m_Multi, 0);
    }
    int DeviceCount() const { return m_Multi ? cndl_multi_device_count(m_Multi) : (m_Ctx ? 1 : 0); }

    // Intersector.h:170-198.  The mesh concatenation of BuildBVH (BVHConstructor.cpp:981-1002) happens here on
    // the host; the build itself runs on the GPU and is byte-identical to BVH::BuildBVH.
    template <typename ObjectT>
    void AddObject(const ObjectT& object, const cndl_build_opts* Options = nullptr) {
        Require();
        std::vector<Vertex> Vertices;
        std::vector<std::uint32_t> MeshIndices;
        std::vector<std::int32_t> MeshReferences;
        detail::ConcatenateMeshes(object, Vertices, MeshIndices, MeshReferences);
        if (m_Multi)
            CheckMulti(cndl_multi_add_object(m_Multi, static_cast<std::uint32_t>(object.GetID()), reinterpret_cast<const cndl_vertex*>(Vertices.data()),
                                             Vertices.size(), MeshIndices.data(), MeshIndices.size(), MeshReferences.data(), Options));
        else
            Check(cndl_add_object(m_Ctx, static_cast<std::uint32_t>(object.GetID()), reinterpret_cast<const cndl_vertex*>(Vertices.data()),
                                  Vertices.size(), MeshIndices.data(), MeshIndices.size(), MeshReferences.data(), Options));
    }

    // Intersector.h:201-216
    template <typename EntityT>
    void PushEntity(const EntityT& entity) {
        Require();
        const std::uint32_t id = static_cast<std::uint32_t>(entity.m_Object->m_ObjectID);
        const int rc = m_Multi ? cndl_multi_push_entity(m_Multi, id, &entity.m_Model[0][0], entity.m_EmissiveAmount, entity.m_TranslucencyAmount)
                               : cndl_push_entity(m_Ctx, id, &entity.m_Model[0][0], entity.m_EmissiveAmount, entity.m_TranslucencyAmount);
        if (rc == CNDL_ERR_UNKNOWN_OBJECT) throw "Trying to push entity whose parent object hasn't been added to global BVH";
        if (m_Multi) CheckMulti(rc); else Check(rc);
    }
    template <typename EntityT>
    void PushEntities(const std::vector<EntityT*>& Entities) {  // Intersector.h:219-224
        for (const auto& e : Entities) PushEntity(*e);
    }
    void BufferEntities() {  // Intersector.h:227-239
        Require();
        if (m_Multi) CheckMulti(cndl_multi_buffer_entities(m_Multi)); else Check(cndl_buffer_entities(m_Ctx));
    }
    void BufferData(bool ClearCPUData) {  // Intersector.h:322-351
        Require();
        if (m_Multi) CheckMulti(cndl_multi_commit(m_Multi)); else Check(cndl_commit(m_Ctx, ClearCPUData ? 1 : 0));
    }

    // One diffuse-GI frame on the device(s), the way DiffuseTrace.glsl:437-518 runs it (camera rays stand in for the G-buffer):
    // camera rays -> closest hits -> SamplesPerPixel cosine-hemisphere rays per pixel -> IntersectRayIgnoreTransparent ->
    // Bounces - 1 further bounces.  Output: cndl_frame_records(Params) records of Params.out_format, row-major (host memory; pinned
    // memory keeps the copy asynchronous).  With several devices the tiles are dealt round-robin and gathered over NVLink.
    void TraceFrame(const cndl_frame_params& Params, void* Output) {
        Require();
        if (m_Multi) CheckMulti(cndl_multi_trace_frame(m_Multi, &Params, Output)); else Check(cndl_trace_frame(m_Ctx, &Params, Output));
    }
    // Two frames in flight: Submit(slot) returns at once, Wait(slot) blocks until that frame's records are in Output.
    void SubmitFrame(const cndl_frame_params& Params, void* Output, int Slot) {
        Require();
        if (m_Multi) CheckMulti(cndl_multi_frame_submit(m_Multi, &Params, Output, Slot)); else Check(cndl_frame_submit(m_Ctx, &Params, Output, Slot));
    }
    void WaitFrame(int Slot) {
        Require();
        if (m_Multi) CheckMulti(cndl_multi_frame_wait(m_Multi, Slot)); else Check(cndl_frame_wait(m_Ctx, Slot));
    }
    // Same shape as IntersectPrimary (Intersector.h:241-266), one bounce further: the compact hit record {t, tri, v, w} of every
    // pixel's diffuse ray(s).
    void IntersectDiffuse(cndl_hit16* Output, int Width, int Height, const float* InverseView, const float* InverseProjection, std::uint32_t Seed,
                          int SamplesPerPixel = 1) {
        cndl_frame_params p;
        std::memset(&p, 0, sizeof(p));
        std::memcpy(p.inv_view, InverseView, 64);
        std::memcpy(p.inv_proj, InverseProjection, 64);
        p.width = Width; p.height = Height; p.spp = SamplesPerPixel; p.bounces = 1; p.seed = Seed;
        p.out_format = CNDL_FRAME_OUT_HIT16;
        p.flags = CNDL_FRAME_OCTANT_ORDER;
        TraceFrame(p, Output);
    }

    // Intersector.h:241-266 with hit records instead of an albedo image. Matrices are glm::mat4-compatible (column-major).
    void IntersectPrimary(RayHit* Output, int Width, int Height, const float* InverseView, const float* InverseProjection) {
        Require();
        Check(cndl_intersect_primary(m_Ctx, InverseView, InverseProjection, Width, Height, reinterpret_cast<cndl_hit*>(Output), nullptr));
    }

    // IntersectRay / IntersectRayIgnoreTransparent of TraverseBVH.glsl for a batch of rays (host buffers).
    void IntersectRays(const Ray* Rays, std::size_t Count, RayHit* Hits, bool IgnoreTransparent = false) {
        Require();
        Check(cndl_intersect_closest(m_Ctx, reinterpret_cast<const cndl_ray*>(Rays), Count, IgnoreTransparent ? CNDL_IGNORE_TRANSPARENT : 0,
                                     reinterpret_cast<cndl_hit*>(Hits)));
    }
    // float IntersectRay(o, d): first accepted t or -1.
    void IntersectRaysAny(const Ray* Rays, std::size_t Count, float* Traversals) {
        Require();
        Check(cndl_intersect_any(m_Ctx, reinterpret_cast<const cndl_ray*>(Rays), Count, Traversals));
    }

    // GetData (Include/TraverseBVHStackless.glsl:370-408) without the texture fetch: interpolated normal / UV, entity emissive / alpha.
    void GetData(const RayHit* Hits, std::size_t Count, HitData* Output) {
        Require();
        Check(cndl_get_data(m_Ctx, reinterpret_cast<const cndl_hit*>(Hits), Count, reinterpret_cast<cndl_hit_attr*>(Output)));
    }

    // GenerateMeshTextureReferences (Intersector.h:367-410): texture-array indices per mesh + upload of the table GetData reads.
    // The reference pulls the materials from FileLoader::GetMeshTexturePaths(); here the caller passes them.  m_TextureHandles[i]
    // is the handle to bind to Textures[i] (m_TextureHandleReferenceMap, :423-427).
    void GenerateMeshTextureReferences(const std::vector<FileLoader::_MeshMaterialData>& MeshMaterials) {
        Require();
        m_MeshTextureReferences.assign(MeshMaterials.size(), BVH::TextureReferences{});
        m_TextureHandles.assign(2 * MeshMaterials.size(), 0);
        std::size_t used = 0;
        Check(cndl_generate_texture_references(reinterpret_cast<const cndl_mesh_material*>(MeshMaterials.data()), MeshMaterials.size(),
                                               reinterpret_cast<cndl_texture_reference*>(m_MeshTextureReferences.data()), m_TextureHandles.data(),
                                               m_TextureHandles.size(), &used));
        m_TextureHandles.resize(used);
        Check(cndl_set_texture_references(m_Ctx, reinterpret_cast<const cndl_texture_reference*>(m_MeshTextureReferences.data()), m_MeshTextureReferences.size()));
    }
    // GetData with the Albedo decision (:393-404): AlbedoRef > -1 = sample Textures[AlbedoRef] at UV, else Albedo = ModelColor.
    void GetData(const RayHit* Hits, std::size_t Count, HitMaterial* Output) {
        Require();
        Check(cndl_get_data_material(m_Ctx, reinterpret_cast<const cndl_hit*>(Hits), Count, reinterpret_cast<cndl_hit_material*>(Output)));
    }

    // Physics::CollideBox / CollidePoint (Physics.cpp:175-228) on the GPU; stackless intersectors only, like Physics.h:15.
    bool CollideBox(const float* Min, const float* Max) {
        Require();
        cndl_box b = {{Min[0], Min[1], Min[2]}, 0.0f, {Max[0], Max[1], Max[2]}, 0.0f};
        cndl_collision c;
        Check(cndl_collide_boxes(m_Ctx, &b, 1, &c));
        return c.collided != 0;
    }
    bool CollidePoint(const float* Point) {
        const float mn[3] = {Point[0] - 0.01f, Point[1] - 0.01f, Point[2] - 0.01f}, mx[3] = {Point[0] + 0.01f, Point[1] + 0.01f, Point[2] + 0.01f};
        return CollideBox(mn, mx);
    }

    // Replacement for BindEverything (Intersector.h:269-320): device pointers of the five buffers.
    void DeviceBuffers(const T** Nodes, const BVH::Triangle** Triangles, const Vertex** Vertices, const BVHEntity** Entities) {
        Require();
        const void* n = nullptr;
        const cndl_triangle* t = nullptr;
        const cndl_vertex* v = nullptr;
        const cndl_entity* e = nullptr;
        Check(cndl_device_buffers(m_Ctx, &n, &t, &v, &e));
        if (Nodes) *Nodes = static_cast<const T*>(n);
        if (Triangles) *Triangles = reinterpret_cast<const BVH::Triangle*>(t);
        if (Vertices) *Vertices = reinterpret_cast<const Vertex*>(v);
        if (Entities) *Entities = reinterpret_cast<const BVHEntity*>(e);
    }

    // Fills the public host arrays the reference exposes (Intersector.h:96-98; read by Physics.cpp:100-126).
    void FetchCPUData() {
        Require();
        m_BVHNodes.resize(cndl_node_count(m_Ctx));
        m_BVHTriangles.resize(cndl_triangle_count(m_Ctx));
        m_BVHVertices.resize(cndl_vertex_count(m_Ctx));
        Check(cndl_read_buffers(m_Ctx, m_BVHNodes.data(), reinterpret_cast<cndl_triangle*>(m_BVHTriangles.data()),
                                reinterpret_cast<cndl_vertex*>(m_BVHVertices.data())));
    }

    bool Collide(const float* Point) { return CollidePoint(Point); }  // Intersector.h:354-358 is a stub in the reference; here it answers
    void Recompile() {}                                      // Intersector.h:361-365
    cndl_ctx* Context() { return m_Ctx; }

    std::vector<T> m_BVHNodes;
    std::vector<Vertex> m_BVHVertices;
    std::vector<BVH::Triangle> m_BVHTriangles;
    std::vector<BVH::TextureReferences> m_MeshTextureReferences;  // Intersector.h:116
    std::vector<std::uint64_t> m_TextureHandles;

private:
    void Require() { if (!m_Ctx) Initialize(0); }
    void Check(int rc) {
        if (rc != CNDL_OK) {
            m_LastError = cndl_last_error(m_Ctx);
            throw m_LastError.c_str();  // the reference throws C strings (Intersector.h:147,:204)
        }
    }
    void CheckMulti(int rc) {
        if (rc != CNDL_OK) {
            m_LastError = cndl_multi_last_error(m_Multi);
            throw m_LastError.c_str();
        }
    }
    cndl_ctx* m_Ctx = nullptr;
    cndl_multi* m_Multi = nullptr;
    bool m_Stackless = false;
    std::string m_LastError;
};

namespace Physics {  // Physics.h:14-15
inline bool CollideBox(const float* Min, const float* Max, RayIntersector<BVH::StacklessTraversalNode>& Intersector) { return Intersector.CollideBox(Min, Max); }
inline bool CollidePoint(const float* Point, RayIntersector<BVH::StacklessTraversalNode>& Intersector) { return Intersector.CollidePoint(Point); }
}  // namespace Physics

}  // namespace Candela
