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
    return cases
