// TEST INFRASTRUCTURE ONLY — see oracle.h.
//
// Restatement of the reference's CPU BVH builder.  Every function names the
// reference lines it follows (paths relative to
// /root/reference/Source/Core/BVH/).  The tree is kept in an index-addressed
// vector instead of the reference's heap nodes; the arithmetic, the order of
// comparisons and the order in which ranges are processed are the reference's.
//
// Where the reference has undefined behaviour the oracle defines it:
//   * fewer than 100 triangles: the reference divides by zero in its progress
//     print (BVHConstructor.cpp:387,:449); the oracle simply builds.
//   * a root with <= 2 triangles in stack format: the reference dereferences
//     null children (:873); the oracle emits one slot whose left child is the
//     leaf and whose right child is an empty leaf (sentinel box, pack 0).
//   * no axis with non-zero extent, or no cost below 1e29: the reference
//     leaves axis/border uninitialised (:276-363); the oracle uses axis 0 and
//     border = node.min[0], which sends the node to the midpoint fallback.
#include "oracle.h"

#include <cstring>
#include <vector>

namespace {

constexpr float kSentinelMax = 10000000.0f;   // BVHConstructor.h:20
constexpr float kSentinelMin = -10000000.0f;  // BVHConstructor.h:21
constexpr int kBins = 64;                      // BVHConstructor.cpp:46
constexpr uint32_t kMaxLeaf = 2;               // BVHConstructor.cpp:50
constexpr float kInfCost = 1e29f;              // BVHConstructor.cpp:56

// glm 0.9.8.5 scalar min/max (glm/detail/func_common.inl:15-28). Argument
// order matters for +0/-0 ties, so call sites keep the reference's order.
inline float gmin(float x, float y) { return x < y ? x : y; }
inline float gmax(float x, float y) { return x > y ? x : y; }

struct Box {
    float mn[3] = {kSentinelMax, kSentinelMax, kSentinelMax};   // Bounds() BVHConstructor.h:27
    float mx[3] = {kSentinelMin, kSentinelMin, kSentinelMin};
    // Bounds::GetArea, BVHConstructor.h:41-44
    float area() const {
        const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
        return ex * ey + ey * ez + ez * ex;
    }
};

struct TreeNode {
    Box box;
    uint32_t start = 0;  // range into refs while building; leaf: sorted position + t_offset
    uint32_t len = 0;    // 0 once the node became inner
    int32_t left = -1, right = -1;
    bool leaf = false;
    bool flip = false;   // children exchanged at flatten time (stackless only)
    uint32_t range_start = 0, range_len = 0;  // build-time range, kept for hashing / flip adoption
};

struct Vertex32 { float pos[4]; uint32_t packed[4]; };  // Utils/Vertex.h:7-12
struct Tri16 { int32_t v[4]; };                          // BVHConstructor.h:79-84
struct Node32 { float mn[4]; float mx[4]; };             // BVHConstructor.h:66-70
struct Node64 { Node32 l, r; };                          // BVHConstructor.h:72-77

inline float bits_to_float(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int32_t float_to_bits(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }

inline uint64_t mix64(uint64_t x) {  // SplitMix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

}  // namespace

struct orc_bvh {
    int format = ORC_STACKLESS;
    int32_t t_offset = 0;
    std::vector<TreeNode> tree;     // tree[0] is the root
    std::vector<int32_t> sorted;    // SortedTriangleReferences
    std::vector<Tri16> tris;
    std::vector<Node32> flat32;
    std::vector<Node64> flat64;
    uint64_t created = 0, leaves = 0, split_fails = 0, max_stack = 0, stack_slots = 0, depth = 0;

    void flatten_stackless();
    void flatten_stack();
};

namespace {

// SearchSAHPlaneBinned, BVHConstructor.cpp:276-363
void binned_sah(const TreeNode& node, const std::vector<int32_t>& refs, const std::vector<Box>& tri_box,
                const std::vector<float>& centroid /* 3 per tri */, int& o_axis, float& o_border) {
    float best = kInfCost;
    o_axis = 0;
    o_border = node.box.mn[0];
    for (int axis = 0; axis < 3; ++axis) {
        const float lo = node.box.mn[axis], hi = node.box.mx[axis];
        if (lo == hi) continue;  // :285
        int count[kBins];
        Box bin[kBins];
        for (int b = 0; b < kBins; ++b) count[b] = 0;
        const float extent = hi - lo;
        const float scale = (float)kBins / extent;  // :295
        for (uint32_t i = node.start; i < node.start + node.len; ++i) {
            const int32_t r = refs[i];
            const float c = centroid[3 * (size_t)r + axis];
            int b = (int)((c - lo) * scale);  // :302
            b = (kBins - 1) < b ? (kBins - 1) : b;
            if (b < 0) b = 0;  // NaN / out-of-range input: undefined in the reference
            count[b]++;
            const Box& tb = tri_box[r];
            for (int k = 0; k < 3; ++k) {
                bin[b].mn[k] = gmin(bin[b].mn[k], tb.mn[k]);  // :306
                bin[b].mx[k] = gmax(bin[b].mx[k], tb.mx[k]);  // :307
            }
        }
        float l_area[kBins - 1], r_area[kBins - 1];
        int l_count[kBins - 1], r_count[kBins - 1];
        Box lbox, rbox;
        int lsum = 0, rsum = 0;
        for (int i = 0; i < kBins - 1; ++i) {  // :319-343
            lsum += count[i];
            l_count[i] = lsum;
            for (int k = 0; k < 3; ++k) {
                lbox.mn[k] = gmin(lbox.mn[k], bin[i].mn[k]);
                lbox.mx[k] = gmax(lbox.mx[k], bin[i].mx[k]);
            }
            l_area[i] = lbox.area();
            const int j = kBins - 1 - i;
            rsum += count[j];
            r_count[j - 1] = rsum;
            for (int k = 0; k < 3; ++k) {
                rbox.mn[k] = gmin(rbox.mn[k], bin[j].mn[k]);
                rbox.mx[k] = gmax(rbox.mx[k], bin[j].mx[k]);
            }
            r_area[j - 1] = rbox.area();
        }
        const float step = extent / (float)kBins;  // :347
        for (int i = 0; i < kBins - 1; ++i) {
            const float cost = (float)l_count[i] * l_area[i] + (float)r_count[i] * r_area[i];  // :351
            if (cost < best) {
                best = cost;
                o_axis = axis;
                o_border = lo + (step * (float)(i + 1));  // :356
            }
        }
    }
}

// ConstructTree, BVHConstructor.cpp:385-627
void construct_tree(orc_bvh& b, const Vertex32* verts, const uint32_t* indices, uint64_t I, int swap_policy,
                    uint64_t swap_seed) {
    const uint32_t T = (uint32_t)(I / 3);
    std::vector<int32_t> refs(T);
    for (uint32_t i = 0; i < T; ++i) refs[i] = (int32_t)i;  // :398-402

    std::vector<Box> tri_box(T);
    std::vector<float> centroid(3 * (size_t)T);
    Box root_box;
    for (uint32_t t = 0; t < T; ++t) {  // :411-427
        Box cur;
        for (int c = 0; c < 3; ++c) {
            const float* p = verts[indices[3 * (size_t)t + c]].pos;
            for (int k = 0; k < 3; ++k) {
                cur.mn[k] = gmin(cur.mn[k], p[k]);
                cur.mx[k] = gmax(cur.mx[k], p[k]);
            }
        }
        for (int k = 0; k < 3; ++k) {
            root_box.mn[k] = gmin(root_box.mn[k], cur.mn[k]);
            root_box.mx[k] = gmax(root_box.mx[k], cur.mx[k]);
            centroid[3 * (size_t)t + k] = (cur.mn[k] + cur.mx[k]) / 2.0f;  // Bounds::GetCenter, BVHConstructor.h:33
        }
        tri_box[t] = cur;
    }

    b.tree.clear();
    b.tree.reserve(2 * (size_t)T + 1);
    TreeNode root;
    root.box = root_box;
    root.start = 0;
    root.len = T;
    root.range_start = 0;
    root.range_len = T;
    b.tree.push_back(root);

    std::vector<int32_t> work;  // LIFO of node ids, :432-443
    work.push_back(0);
    b.sorted.clear();
    b.sorted.reserve(T);

    while (!work.empty()) {
        if (work.size() > b.max_stack) b.max_stack = work.size();  // :445
        const int32_t id = work.back();
        work.pop_back();

        if (b.tree[id].len <= kMaxLeaf || b.tree[id].len <= 1) {  // :456
            TreeNode& n = b.tree[id];
            n.leaf = true;
            b.leaves++;
            const int32_t at = (int32_t)b.sorted.size();
            for (uint32_t i = n.start; i < n.start + n.len; ++i) b.sorted.push_back(refs[i]);
            n.start = (uint32_t)(at + b.t_offset);  // :469
            continue;
        }

        int axis;
        float border;
        binned_sah(b.tree[id], refs, tri_box, centroid, axis, border);  // :481

        const uint32_t start = b.tree[id].start, len = b.tree[id].len;
        uint32_t mid = start;
        for (uint32_t i = start; i < start + len; ++i) {  // :532-549
            if (centroid[3 * (size_t)refs[i] + axis] < border) {
                const int32_t tmp = refs[i];
                refs[i] = refs[mid];
                refs[mid] = tmp;
                ++mid;
            }
        }
        if (mid == start || mid == start + len) {  // :553-556
            mid = start + len / 2;
            b.split_fails++;
        }

        TreeNode l, r;
        b.created += 2;  // LastNodeIndex++, twice (:559,:566)
        l.start = l.range_start = start;
        l.len = l.range_len = mid - start;
        r.start = r.range_start = mid;
        r.len = r.range_len = (start + len) - mid;
        for (uint32_t x = 0; x < l.len; ++x) {  // :577-583
            const Box& tb = tri_box[refs[l.start + x]];
            for (int k = 0; k < 3; ++k) {
                l.box.mn[k] = gmin(tb.mn[k], l.box.mn[k]);
                l.box.mx[k] = gmax(tb.mx[k], l.box.mx[k]);
            }
        }
        for (uint32_t x = 0; x < r.len; ++x) {  // :589-593
            const Box& tb = tri_box[refs[r.start + x]];
            for (int k = 0; k < 3; ++k) {
                r.box.mn[k] = gmin(tb.mn[k], r.box.mn[k]);
                r.box.mx[k] = gmax(tb.mx[k], r.box.mx[k]);
            }
        }

        const int32_t li = (int32_t)b.tree.size();
        b.tree.push_back(l);
        b.tree.push_back(r);
        TreeNode& n = b.tree[id];
        n.len = 0;  // :612
        n.left = li;
        n.right = li + 1;
        // :599-609 — the label flip never changes which range is processed first.
        if (b.format == ORC_STACKLESS && swap_policy == ORC_SWAP_HASHED) {
            n.flip = (mix64(swap_seed ^ mix64(((uint64_t)start << 32) | len)) & 1ull) != 0;
        }
        work.push_back(li);      // left pushed first ...
        work.push_back(li + 1);  // ... so the right range is popped first (:624-625)
    }

    // GenerateTriangles, :630-650
    b.tris.resize(b.sorted.size());
    for (size_t i = 0; i < b.sorted.size(); ++i) {
        const size_t r = (size_t)b.sorted[i];
        b.tris[i].v[0] = (int32_t)indices[3 * r + 0];
        b.tris[i].v[1] = (int32_t)indices[3 * r + 1];
        b.tris[i].v[2] = (int32_t)indices[3 * r + 2];
        b.tris[i].v[3] = 0;  // mesh id filled by the caller
    }
}

inline int32_t leaf_pack(const TreeNode& n) { return (int32_t)((n.start << 4) | (n.len & 0xF)); }  // :794

}  // namespace

// FlattenBVH + FlattenBVHRecursive, BVHConstructor.cpp:783-845
void orc_bvh::flatten_stackless() {
    const size_t total = (size_t)created + 1;
    flat32.assign(total, Node32{});
    std::vector<int32_t> right_of(total, -1), pack_of(total, 0);
    // Pre-order numbering, first child first; explicit stack instead of recursion.
    struct Frame { int32_t node; int32_t parent_slot; };
    std::vector<Frame> st;
    st.push_back({0, -1});
    int32_t next = 0;
    uint64_t max_depth = 0;
    std::vector<uint32_t> depth_of(total, 0);
    while (!st.empty()) {
        const Frame f = st.back();
        st.pop_back();
        const TreeNode& n = tree[f.node];
        const int32_t slot = next++;
        if (f.parent_slot >= 0) {
            right_of[f.parent_slot] = slot;  // this frame is a second child
            depth_of[slot] = depth_of[f.parent_slot] + 1;
        } else if (slot > 0) {
            depth_of[slot] = depth_of[slot - 1] + 1;  // first child follows its parent
        }
        if (depth_of[slot] > max_depth) max_depth = depth_of[slot];
        for (int k = 0; k < 3; ++k) {
            flat32[slot].mn[k] = n.box.mn[k];
            flat32[slot].mx[k] = n.box.mx[k];
        }
        flat32[slot].mn[3] = 0.0f;
        flat32[slot].mx[3] = 0.0f;
        if (n.leaf) {
            pack_of[slot] = leaf_pack(n);
            right_of[slot] = -1;
        } else {
            const int32_t first = n.flip ? n.right : n.left;
            const int32_t second = n.flip ? n.left : n.right;
            right_of[slot] = -2;                // inner; patched when `second` is numbered
            st.push_back({second, slot});
            st.push_back({first, -1});
        }
    }
    depth = max_depth;
    // Links, :813-830
    flat32[0].mx[3] = bits_to_float(-1);
    for (size_t i = 0; i < total; ++i) {
        if (!tree.empty() && right_of[i] >= 0) {
            flat32[i + 1].mx[3] = bits_to_float(right_of[i]);
            flat32[right_of[i]].mx[3] = flat32[i].mx[3];
        }
    }
    // Leaf packs / inner flag, :834-844
    for (size_t i = 0; i < total; ++i) flat32[i].mn[3] = bits_to_float(right_of[i] >= 0 ? -1 : pack_of[i]);
}

// FlattenStackBVH, BVHConstructor.cpp:847-930
void orc_bvh::flatten_stack() {
    const size_t total = (size_t)created + 1;
    flat64.assign(total, Node64{});  // value-initialised: unused slots are all-zero (:762-765)
    stack_slots = 0;
    auto fill_child = [&](Node32& out, const TreeNode& c) {
        for (int k = 0; k < 3; ++k) { out.mn[k] = c.box.mn[k]; out.mx[k] = c.box.mx[k]; }
        out.mn[3] = 0.0f;
        out.mx[3] = 0.0f;
        out.mn[3] = bits_to_float(c.leaf ? leaf_pack(c) : -1);
    };
    if (tree[0].leaf) {  // defined here; the reference crashes
        fill_child(flat64[0].l, tree[0]);
        TreeNode empty;
        empty.leaf = true;
        fill_child(flat64[0].r, empty);
        stack_slots = 1;
        depth = 0;
        return;
    }
    struct Item { int32_t node; int32_t tag; uint32_t depth; };  // tag: +slot+1 left of slot, -(slot+1) right of slot
    std::vector<Item> q;
    q.push_back({0, 0, 0});
    size_t head = 0;
    int32_t counter = 0;
    uint64_t max_depth = 0;
    while (head < q.size()) {
        const Item it = q[head++];
        const TreeNode& n = tree[it.node];
        Node64& slot = flat64[counter++];
        const TreeNode& lc = tree[n.left];
        const TreeNode& rc = tree[n.right];
        if (it.depth + 1 > max_depth) max_depth = it.depth + 1;
        fill_child(slot.l, lc);
        if (!lc.leaf) q.push_back({n.left, counter, it.depth + 1});
        fill_child(slot.r, rc);
        if (!rc.leaf) q.push_back({n.right, -counter, it.depth + 1});
        if (it.tag > 0) flat64[it.tag - 1].l.mx[3] = bits_to_float(counter - 1);
        else if (it.tag < 0) flat64[-it.tag - 1].r.mx[3] = bits_to_float(counter - 1);
    }
    stack_slots = (uint64_t)counter;
    depth = max_depth;
}

extern "C" {

orc_bvh* orc_build(int format, const void* verts, uint64_t V, const uint32_t* indices, uint64_t I,
                   const int32_t* mesh_id_per_tri, int32_t t_offset, int swap_policy, uint64_t swap_seed) {
    if (!verts || !indices || I == 0 || I % 3 != 0) return nullptr;
    if (format != ORC_STACKLESS && format != ORC_STACK) return nullptr;
    for (uint64_t i = 0; i < I; ++i)
        if (indices[i] >= V) return nullptr;
    orc_bvh* b = new orc_bvh;
    b->format = format;
    b->t_offset = t_offset;
    construct_tree(*b, static_cast<const Vertex32*>(verts), indices, I, swap_policy, swap_seed);
    for (size_t i = 0; i < b->tris.size(); ++i)
        b->tris[i].v[3] = mesh_id_per_tri ? mesh_id_per_tri[b->sorted[i]] : 0;
    if (format == ORC_STACKLESS) b->flatten_stackless();
    else b->flatten_stack();
    return b;
}

void orc_bvh_free(orc_bvh* b) { delete b; }
uint64_t orc_bvh_node_count(const orc_bvh* b) { return b->created + 1; }
uint64_t orc_bvh_tri_count(const orc_bvh* b) { return b->tris.size(); }

void orc_bvh_stats(const orc_bvh* b, uint64_t s[6]) {
    s[0] = b->created; s[1] = b->leaves; s[2] = b->split_fails; s[3] = b->max_stack; s[4] = b->stack_slots; s[5] = b->depth;
}

void orc_bvh_fetch(const orc_bvh* b, void* nodes, void* tris) {
    if (nodes) {
        if (b->format == ORC_STACKLESS) std::memcpy(nodes, b->flat32.data(), b->flat32.size() * sizeof(Node32));
        else std::memcpy(nodes, b->flat64.data(), b->flat64.size() * sizeof(Node64));
    }
    if (tris) std::memcpy(tris, b->tris.data(), b->tris.size() * sizeof(Tri16));
}

void orc_bvh_fetch_order(const orc_bvh* b, int32_t* out) { std::memcpy(out, b->sorted.data(), b->sorted.size() * 4); }

int64_t orc_bvh_adopt_flips(orc_bvh* b, const void* ref_nodes, uint64_t n_nodes) {
    if (b->format != ORC_STACKLESS || n_nodes != b->created + 1) return -1;
    const Node32* ref = static_cast<const Node32*>(ref_nodes);
    const uint32_t T = (uint32_t)b->sorted.size();
    // Sorted range of a tree node: leaves are emitted right range first, so the
    // build range [s, s+len) lands at sorted positions [T-(s+len), T-s).
    auto sorted_lo = [&](const TreeNode& n) { return T - (n.range_start + n.range_len); };
    int64_t flips = 0;
    struct Pair { int32_t node; int64_t slot; };
    std::vector<Pair> st;
    st.push_back({0, 0});
    while (!st.empty()) {
        const Pair p = st.back();
        st.pop_back();
        TreeNode& n = b->tree[p.node];
        if ((uint64_t)p.slot >= n_nodes) return -1;
        const bool ref_inner = float_to_bits(ref[p.slot].mn[3]) == -1;
        if (n.leaf) {
            if (ref_inner) return -1;
            continue;
        }
        if (!ref_inner) return -1;
        // first triangle under the reference's first child: follow slot+1 down to a leaf
        int64_t s = p.slot + 1;
        while ((uint64_t)s < n_nodes && float_to_bits(ref[s].mn[3]) == -1) ++s;
        if ((uint64_t)s >= n_nodes) return -1;
        const uint32_t first_tri = (uint32_t)(float_to_bits(ref[s].mn[3]) >> 4) - (uint32_t)b->t_offset;
        const TreeNode& l = b->tree[n.left];
        const uint32_t llo = sorted_lo(l), lhi = llo + l.range_len;
        n.flip = !(first_tri >= llo && first_tri < lhi);
        if (n.flip) ++flips;
        const int64_t second_slot = float_to_bits(ref[p.slot + 1].mx[3]);
        if (second_slot < 0) return -1;
        st.push_back({n.flip ? n.left : n.right, second_slot});
        st.push_back({n.flip ? n.right : n.left, p.slot + 1});
    }
    b->flatten_stackless();
    return flips;
}

double orc_bvh_sah_cost(const orc_bvh* b) {
    const double root_area = (double)b->tree[0].box.area();
    if (!(root_area > 0.0)) return 0.0;
    double cost = 0.0;
    for (const TreeNode& n : b->tree) {
        const double a = (double)n.box.area() / root_area;
        cost += n.leaf ? a * (double)n.len : a;
    }
    return cost;
}

void orc_concat_meshes(int n_meshes, const void* verts, const uint64_t* vcount, const uint32_t* indices,
                       const uint64_t* icount, const int32_t* mesh_numbers, void* out_verts, uint32_t* out_indices,
                       int32_t* out_mesh_ids) {
    // BVHConstructor.cpp:981-1002
    const Vertex32* vin = static_cast<const Vertex32*>(verts);
    Vertex32* vout = static_cast<Vertex32*>(out_verts);
    uint32_t index_offset = 0;
    uint64_t vo = 0, io = 0, to = 0;
    for (int m = 0; m < n_meshes; ++m) {
        for (uint64_t x = 0; x < icount[m]; ++x) {
            out_indices[io + x] = indices[io + x] + index_offset;
            if (x % 3 == 0) out_mesh_ids[to++] = mesh_numbers[m];
        }
        for (uint64_t x = 0; x < vcount[m]; ++x) vout[vo + x] = vin[vo + x];
        io += icount[m];
        vo += vcount[m];
        index_offset += (uint32_t)vcount[m];
    }
}

void orc_rebase_triangles(void* tris, uint64_t T, uint32_t index_offset) {
    // Intersector.h:190-197
    Tri16* t = static_cast<Tri16*>(tris);
    for (uint64_t i = 0; i < T; ++i) {
        t[i].v[0] += (int32_t)index_offset;
        t[i].v[1] += (int32_t)index_offset;
        t[i].v[2] += (int32_t)index_offset;
    }
}

}  // extern "C"
