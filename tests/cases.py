"""Traversal test scenes shared by the CPU (oracle) and GPU (parity) tests.  Each case is built with
the oracle (test infrastructure) and handed to the GPU path as prebuilt reference-layout buffers."""
import numpy as np

from helpers import rays_in_box


def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


def scale_rot(s, deg, t=(0, 0, 0)):
    c, si = np.cos(np.deg2rad(deg)), np.sin(np.deg2rad(deg))
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = np.array([[c, 0, si], [0, 1, 0], [-si, 0, c]], np.float32) * np.float32(s)
    m[:3, 3] = t
    return m


def two_triangles():
    """Triangle 1 of the sorted buffer is the known-answer target; triangle 0 is the blind one."""
    P = np.array([[-1, -1, 5], [1, -1, 5], [0, 1, 5], [9, 9, 9], [10, 9, 9], [9, 10, 9]], np.float32)
    F = np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
    return P, F


def long_corridor(n=1500):
    """n thin slabs along +x whose boxes a ray along x enters but whose triangles it misses: a walk
    longer than the 1024-iteration cap."""
    i = np.arange(n, dtype=np.float32)
    A = np.stack([i, np.full(n, 0.5, np.float32), np.full(n, -1, np.float32)], 1)
    B = np.stack([i + 0.5, np.full(n, 0.5, np.float32), np.full(n, -1, np.float32)], 1)
    C = np.stack([i, np.full(n, -0.5, np.float32), np.full(n, 1, np.float32)], 1)
    P = np.stack([A, B, C], 1).reshape(-1, 3)
    return P.astype(np.float32), np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)


def build_cases(ob, golden_meshes, fmt):
    """Returns a list of dicts {name, scene (ob.Scene), rays}."""
    cases = []

    def add(name, sc, rays):
        cases.append(dict(name=name, scene=sc, rays=np.ascontiguousarray(rays, dtype=ob.RAY_DT)))

    # dragon, random rays from inside the box
    P, F = golden_meshes["dragon"]
    sc = ob.Scene(fmt)
    sc.add_object(2, ob.make_vertices(P), F.ravel(), np.full(len(F), 7, np.int32))
    sc.push_entity(2)
    add("dragon_random", sc, rays_in_box(P.min(0), P.max(0), 20000, 1))

    # axis-parallel rays and origins exactly on box planes: the 0*inf = NaN slab path
    lo, hi = P.min(0), P.max(0)
    r = rays_in_box(lo, hi, 6000, 2)
    r["d"][:2000] = (1, 0, 0)
    r["d"][2000:4000] = (0, -1, 0)
    r["d"][4000:] = (0, 0, 1)
    r["o"][::3, 1] = lo[1]
    r["o"][1::3, 2] = hi[2]
    r["o"][2::3, 0] = np.float32(0.5) * (lo[0] + hi[0])
    add("dragon_axis_parallel", sc, r)

    # several objects and instanced / scaled / translucent entities
    sc2 = ob.Scene(fmt)
    Pc, Fc = golden_meshes["peach_castle"]
    Ps, Fs = golden_meshes["soup400"]
    sc2.add_object(2, ob.make_vertices(Pc), Fc.ravel(), np.full(len(Fc), 1, np.int32), swap_policy=ob.SWAP_HASHED, swap_seed=5)
    sc2.add_object(3, ob.make_vertices(Ps), Fs.ravel(), np.full(len(Fs), 2, np.int32))
    sc2.add_object(4, ob.make_vertices(P), F.ravel(), np.full(len(F), 3, np.int32))
    ext = float(np.abs(Pc).max())
    sc2.push_entity(3, model=translate(0.3, 0.1, -0.2))
    sc2.push_entity(2, model=scale_rot(6.0 / ext, 30.0))
    sc2.push_entity(4, model=scale_rot(0.4, -70.0, (1.0, -1.0, 0.5)), translucency=0.5)
    sc2.push_entity(3, model=scale_rot(1.7, 120.0, (-2.0, 0.5, 1.0)), emissive=2.0)
    sc2.push_entity(2, model=scale_rot(3.0 / ext, 200.0, (0.5, 2.0, 0.0)), translucency=0.005)
    add("multi_entity", sc2, rays_in_box((-7, -7, -7), (7, 7, 7), 20000, 3))

    # known answers + the global-triangle-0 blind spot
    Pt, Ft = two_triangles()
    sc3 = ob.Scene(fmt)
    sc3.add_object(2, ob.make_vertices(Pt), Ft.ravel(), np.array([11, 12], np.int32))
    sc3.push_entity(2)
    kr = np.zeros(4, dtype=ob.RAY_DT)
    kr["o"] = [(0, 0, 0), (9.25, 9.25, 0), (0, 0, 0), (0, 0, 6)]
    kr["d"] = [(0, 0, 1), (0, 0, 1), (0, 0, -1), (0, 0, 1)]
    add("known_answers", sc3, kr)

    # walk longer than the iteration cap
    Pl, Fl = long_corridor()
    sc4 = ob.Scene(fmt)
    sc4.add_object(2, ob.make_vertices(Pl), Fl.ravel(), None)
    sc4.push_entity(2)
    lr = np.zeros(8, dtype=ob.RAY_DT)
    lr["o"] = [(-1, 0.45 - 0.1 * k, 0.05 * k) for k in range(8)]
    lr["d"] = (1, 0, 0)
    add("iteration_cap", sc4, lr)

    # stack format only: a hand-made chain of slots whose two children are both inner, both hit and point at the
    # same next slot, so every iteration pushes: the walk must stop at StackPointer >= 63 (…Stack.glsl:293-295)
    if fmt == ob.STACK:
        Pt2, Ft2 = two_triangles()
        tmp = ob.Scene(ob.STACK)
        tmp.add_object(2, ob.make_vertices(Pt2), Ft2.ravel(), np.array([11, 12], np.int32))
        n_slots = 80
        nodes = np.zeros(n_slots, dtype=ob.NODE64_DT)
        for k in range(n_slots - 1):
            for side in ("l", "r"):
                nodes[side + "min"][k] = (-2, -2, -1, 0)
                nodes[side + "max"][k] = (2, 2, 20, 0)
                nodes[side + "min"][k, 3:].view(np.int32)[0] = -1
                nodes[side + "max"][k, 3:].view(np.int32)[0] = k + 1
        leaf = tmp.nodes[0]                                   # the real (single-slot) tree: both children are leaves
        nodes[n_slots - 1] = leaf
        sc5 = ob.Scene(ob.STACK)
        sc5.nodes, sc5.tris, sc5.verts = nodes, tmp.tris, tmp.verts
        sc5.objects[2] = dict(node_offset=0, node_count=n_slots, tri_offset=0, tri_count=len(tmp.tris), vert_offset=0, vert_count=len(tmp.verts))
        sc5.builds[2] = tmp.builds[2]
        sc5.push_entity(2)
        kr2 = np.zeros(3, dtype=ob.RAY_DT)
        kr2["o"] = [(0, 0, 0), (0.1, 0.1, 0), (5, 5, 0)]
        kr2["d"] = (0, 0, 1)
        add("stack_overflow", sc5, kr2)

    # degenerate geometry
    for nm in ("coplanar_grid", "duplicates", "collinear", "signed_zero"):
        Pd, Fd = golden_meshes[nm]
        s = ob.Scene(fmt)
        s.add_object(2, ob.make_vertices(Pd), Fd.ravel(), None)
        s.push_entity(2)
        lo, hi = Pd.min(0) - 1, Pd.max(0) + 1
        add(nm, s, rays_in_box(lo, hi, 3000, 4))
    # ties (SURVEY.md §8c "tie on a shared edge"): the 12 x 12 integer grid in the plane z = 0, rays through its vertices (shared by up
    # to six triangles), edge midpoints and diagonal midpoints.  Every coordinate is a small integer or a half, so every product is
    # exact: two to six triangles report the SAME t with u, v of exactly 0, 0.5 or 1, RayTriangle accepts all of them (its test is
    # u < 0 || v < 0 || u + v > 1) and the strict t < TMax keeps the first one the walk meets.
    Pg, Fg = golden_meshes["coplanar_grid"]
    sg = ob.Scene(fmt)
    sg.add_object(2, ob.make_vertices(Pg), Fg.ravel(), (np.arange(len(Fg)) % 4).astype(np.int32))
    sg.push_entity(2)
    k = np.arange(13, dtype=np.float32)
    h = np.arange(12, dtype=np.float32) + np.float32(0.5)
    pts = np.concatenate([np.stack(np.meshgrid(a, b, indexing="ij"), -1).reshape(-1, 2) for a, b in ((k, k), (h, k), (k, h), (h, h))]).astype(np.float32)
    tr = np.zeros(3 * len(pts), dtype=ob.RAY_DT)
    n = len(pts)
    tr["o"][:n] = np.concatenate([pts, np.full((n, 1), 3, np.float32)], 1)                 # straight down
    tr["d"][:n] = (0, 0, -1)
    tr["o"][n:2 * n] = np.concatenate([pts, np.full((n, 1), -2, np.float32)], 1)           # straight up, from below
    tr["d"][n:2 * n] = (0, 0, 1)
    tr["o"][2 * n:] = np.concatenate([pts - np.float32([1.5, 0.5]), np.full((n, 1), 1, np.float32)], 1)   # oblique, unnormalised, still exact: lands on pts
    tr["d"][2 * n:] = (1.5, 0.5, -1)
    tr["tmax"] = 1.0e6
    add("shared_edges", sg, tr)
    return cases


def random_mesh(kind, T, seed):
    """Seeded meshes whose character stresses a different rule of the builder each: ties in the binning, zero-extent axes,
    split failures, long thin boxes, clustered centroids."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-10, 10, size=(T, 1, 3))
    if kind == "soup":
        P = c + rng.normal(scale=0.3, size=(T, 3, 3))
    elif kind == "quantised":          # coordinates on a coarse lattice: many equal centroids and bin-edge ties
        P = np.round(c + rng.normal(scale=1.0, size=(T, 3, 3)))
    elif kind == "flat":               # zero extent on y: the axis is skipped (BVHConstructor.cpp:285)
        P = c + rng.normal(scale=0.5, size=(T, 3, 3))
        P[..., 1] = 2.5
    elif kind == "slivers":            # long thin triangles: boxes far larger than the centroid spread
        P = c + rng.normal(scale=0.05, size=(T, 3, 3))
        P[:, 2, :] += rng.normal(scale=8.0, size=(T, 3))
    elif kind == "clusters":           # a few tight clusters far apart: most bins empty
        k = rng.integers(0, 5, size=T)
        P = (np.array([[-9, 0, 0], [9, 0, 0], [0, 9, 0], [0, 0, -9], [4, 4, 4]], np.float64)[k])[:, None, :] + rng.normal(scale=0.02, size=(T, 3, 3))
    elif kind == "repeats":            # every triangle four times: split failures down the tree
        base = c[: (T + 3) // 4] + rng.normal(scale=0.3, size=((T + 3) // 4, 3, 3))
        P = np.repeat(base, 4, axis=0)[:T]
    else:
        raise ValueError(kind)
    P = P.reshape(-1, 3).astype(np.float32)
    F = np.arange(3 * T, dtype=np.uint32).reshape(-1, 3)
    return P, F


RANDOM_MESH_KINDS = ["soup", "quantised", "flat", "slivers", "clusters", "repeats"]


def random_model(rng):
    """A random affine instance matrix: rotation about a random axis, non-uniform scale (a factor may be negative: a mirrored
    instance), shear now and then, translation."""
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    th = rng.uniform(0, 2 * np.pi)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    S = np.diag(rng.uniform(0.3, 3.0, size=3) * rng.choice([1.0, 1.0, 1.0, -1.0], size=3))
    if rng.random() < 0.3:
        S[0, 1] = rng.uniform(-0.5, 0.5)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = (R @ S).astype(np.float32)
    m[:3, 3] = rng.uniform(-4, 4, size=3).astype(np.float32)
    return m


def random_scene(ob, fmt, seed):
    """(scene, rays): 1-3 objects of random_mesh kinds (hashed child flips on some), 1-5 entities with random_model matrices,
    translucency and emissive flags; 3000 rays from inside the scene's neighbourhood + 1000 aimed at it from far outside, every
    seventh direction scaled by 2.5 (the shaders never renormalise, …Stackless.glsl:177-178)."""
    from helpers import rays_in_box
    rng = np.random.default_rng(100 + seed)
    sc = ob.Scene(fmt)
    n_obj = 1 + seed % 3
    for k in range(n_obj):
        P, F = random_mesh(RANDOM_MESH_KINDS[(seed + k) % len(RANDOM_MESH_KINDS)], int(rng.integers(100, 1500)), 31 * seed + k)
        sc.add_object(2 + k, ob.make_vertices(P * np.float32(0.3)), F.ravel(), np.full(len(F), k + 1, np.int32),
                      swap_policy=ob.SWAP_HASHED if (seed + k) % 2 else ob.SWAP_NONE, swap_seed=seed)
    for e in range(1 + 2 * (seed % 3)):
        sc.push_entity(2 + int(rng.integers(0, n_obj)), model=random_model(rng) if e or seed % 2 else None,
                       emissive=float(rng.random() < 0.3) * 2.0, translucency=float(rng.choice([0.0, 0.0, 0.005, 0.5])))
    rays = rays_in_box((-6, -6, -6), (6, 6, 6), 3000, 50 + seed)
    far = rays_in_box((-40, -40, -40), (40, 40, 40), 1000, 70 + seed)
    far["d"] = -far["o"] / np.linalg.norm(far["o"], axis=1, keepdims=True) + rng.normal(scale=0.1, size=(1000, 3)).astype(np.float32)
    rays = np.concatenate([rays, far])
    rays["d"][::7] *= np.float32(2.5)
    return sc, np.ascontiguousarray(rays)


def random_scene_queries(ob, seed):
    """(kind, tmax) pairs of the sweep; the short any-hit range varies with the seed."""
    t = float(np.random.default_rng(seed).uniform(0.5, 5.0))
    return ((ob.CLOSEST, 0.0), (ob.CLOSEST_IGNORE_TRANSPARENT, 0.0), (ob.ANY, 0.0), (ob.ANY, t))
