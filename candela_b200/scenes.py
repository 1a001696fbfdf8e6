"""Deterministic synthetic scenes and ray batches for the workloads BASELINE.json names.

Everything here is plain numpy in float32 and seeded; the arrays it returns are INPUTS that are
handed unchanged to the GPU path, the CPU oracle and (where it exists) the reference builder, so
no result depends on how they were computed.  Shapes follow SURVEY.md §8(d).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .api import RAY_DT, VERTEX_DT, make_vertices

DATA = Path(__file__).resolve().parent / "data"
F32 = np.float32


# ------------------------------------------------------------------------------------------------
# geometry
def load_dragon():
    """The reference's bundled Source/Models/dragon/dragon.obj (12,500 vertices, 25,000 triangles) as
    converted by tests/golden/make_golden.py. Returns (positions[V,3] f32, faces[T,3] u32)."""
    z = np.load(DATA / "dragon_25k.npz")
    return z["positions"].astype(F32), z["faces"].astype(np.uint32)


def _grid_quad(origin, du, dv, nu, nv):
    """nu x nv quads spanning origin + i*du + j*dv, two triangles each."""
    origin, du, dv = (np.asarray(a, dtype=F32) for a in (origin, du, dv))
    i, j = np.meshgrid(np.arange(nu + 1, dtype=F32), np.arange(nv + 1, dtype=F32), indexing="ij")
    pos = origin[None, None, :] + i[..., None] * du[None, None, :] + j[..., None] * dv[None, None, :]
    pos = pos.reshape(-1, 3).astype(F32)
    a = (np.arange(nu)[:, None] * (nv + 1) + np.arange(nv)[None, :]).ravel()
    faces = np.concatenate([np.stack([a, a + nv + 1, a + nv + 2], 1), np.stack([a, a + nv + 2, a + 1], 1)]).astype(np.uint32)
    return pos, faces


def _cylinder(center_xz, y0, y1, radius, segs, rings):
    th = (np.arange(segs, dtype=F32) * F32(2.0 * np.pi / segs)).astype(F32)
    ys = np.linspace(y0, y1, rings + 1, dtype=F32)
    x = (center_xz[0] + radius * np.cos(th)).astype(F32)
    z = (center_xz[1] + radius * np.sin(th)).astype(F32)
    pos = np.stack([np.tile(x, rings + 1), np.repeat(ys, segs), np.tile(z, rings + 1)], 1).astype(F32)
    faces = []
    for r in range(rings):
        a = r * segs + np.arange(segs)
        b = r * segs + (np.arange(segs) + 1) % segs
        faces.append(np.stack([a, b, b + segs], 1))
        faces.append(np.stack([a, b + segs, a + segs], 1))
    return pos, np.concatenate(faces).astype(np.uint32)


def _arch(x0, x1, z, y_spring, thickness, segs):
    """Half-ring of quads between two columns (an arcade arch), in the plane z = const, extruded in z."""
    cx, r = F32(0.5) * (x0 + x1), F32(0.5) * (x1 - x0)
    th = np.linspace(0.0, np.pi, segs + 1, dtype=F32)
    xs = (cx - r * np.cos(th)).astype(F32)
    ys = (y_spring + r * np.sin(th)).astype(F32)
    front = np.stack([xs, ys, np.full_like(xs, z - thickness)], 1)
    back = np.stack([xs, ys, np.full_like(xs, z + thickness)], 1)
    pos = np.concatenate([front, back]).astype(F32)
    a = np.arange(segs)
    n = segs + 1
    faces = np.concatenate([np.stack([a, a + 1, a + 1 + n], 1), np.stack([a, a + 1 + n, a + n], 1)]).astype(np.uint32)
    return pos, faces


def _rot_y(deg):
    c, s = np.cos(np.deg2rad(deg)), np.sin(np.deg2rad(deg))
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=F32)


def make_s260k():
    """Stand-in for the missing ~260k-triangle Sponza: a closed 40 x 14 x 18 hall with a
    tessellated shell, two colonnades with arches, and ten copies of the reference's dragon mesh baked in
    at fixed transforms.  One object, several meshes.  Returns (verts[VERTEX_DT], indices u32, mesh_ids i32)."""
    parts = []  # (positions, faces, mesh id)
    L, Hh, Wd = F32(40.0), F32(14.0), F32(18.0)
    x0, z0 = -L / 2, -Wd / 2
    parts.append((*_grid_quad([x0, 0, z0], [40.0 / 48.0, 0, 0], [0, 0, 0.75], 48, 24), 0))     # floor
    parts.append((*_grid_quad([x0, Hh, z0], [0, 0, 0.75], [40.0 / 48.0, 0, 0], 24, 48), 1))    # ceiling
    parts.append((*_grid_quad([x0, 0, z0], [0, 1.0, 0], [1.0, 0, 0], 14, 40), 2))              # wall z = -9
    parts.append((*_grid_quad([x0, 0, -z0], [1.0, 0, 0], [0, 1.0, 0], 40, 14), 2))             # wall z = +9
    parts.append((*_grid_quad([x0, 0, z0], [0, 0, 1.0], [0, 1.0, 0], 18, 14), 3))              # wall x = -20
    parts.append((*_grid_quad([-x0, 0, z0], [0, 1.0, 0], [0, 0, 1.0], 14, 18), 3))             # wall x = +20
    col_x = np.linspace(-17.5, 17.5, 8, dtype=F32)
    for zc in (F32(-6.0), F32(6.0)):
        for k, xc in enumerate(col_x):
            parts.append((*_cylinder((xc, zc), F32(0.0), F32(8.0), F32(0.45), 16, 8), 4))
            if k + 1 < len(col_x):
                parts.append((*_arch(xc, col_x[k + 1], zc, F32(8.0), F32(0.3), 24), 5))
    dp, df = load_dragon()
    centre = F32(0.5) * (dp.min(0) + dp.max(0))
    base = dp - np.array([centre[0], dp[:, 1].min(), centre[2]], dtype=F32)
    k = 0
    for zc in (F32(-2.5), F32(2.5)):
        for xc in np.linspace(-14.0, 14.0, 5, dtype=F32):
            m = _rot_y(36.0 * k + 10.0) * F32(0.3 + 0.02 * (k % 3))
            p = (base @ m.T).astype(F32) + np.array([xc, 0.0, zc], dtype=F32)
            parts.append((p.astype(F32), df, 6 + k))
            k += 1
    pos, idx, mid, off = [], [], [], 0
    for p, f, m in parts:
        pos.append(p.astype(F32))
        idx.append(f.astype(np.uint32) + np.uint32(off))
        mid.append(np.full(len(f), m, dtype=np.int32))
        off += len(p)
    return make_vertices(np.concatenate(pos)), np.concatenate(idx).astype(np.uint32).ravel(), np.concatenate(mid)


def make_heightfield(n: int, seed: int = 7):
    """(n-1)^2*2 triangles: an n x n grid over [-50,50]^2 displaced by a few octaves of seeded sines.
    n = 2237 gives 9,999,392 triangles (the ~10M scene of BASELINE config 5)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-50.0, 50.0, n, dtype=F32)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    Y = np.zeros_like(X)
    for o in range(6):
        f = F32(0.08 * 2.0 ** o)
        ph = rng.uniform(0, 2 * np.pi, size=4).astype(F32)
        amp = F32(6.0 / 2.0 ** o)
        Y += amp * (np.sin(f * X + ph[0]) * np.cos(f * Z + ph[1]) + F32(0.5) * np.sin(f * (X + Z) * F32(0.7) + ph[2])).astype(F32)
    pos = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(F32)
    a = (np.arange(n - 1, dtype=np.int64)[:, None] * n + np.arange(n - 1, dtype=np.int64)[None, :]).ravel()
    faces = np.concatenate([np.stack([a, a + n, a + n + 1], 1), np.stack([a, a + n + 1, a + 1], 1)]).astype(np.uint32)
    return make_vertices(pos), faces.ravel(), np.zeros(len(faces), dtype=np.int32)


def make_soup(n_tris: int, seed: int = 3, extent: float = 10.0, size: float = 0.3):
    """Random small triangles in a cube, unshared vertices: the incoherent worst case and a fuzz input."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, size=(n_tris, 1, 3)).astype(F32)
    p = (c + rng.normal(0, size, size=(n_tris, 3, 3)).astype(F32)).reshape(-1, 3).astype(F32)
    return make_vertices(p), np.arange(3 * n_tris, dtype=np.uint32), np.zeros(n_tris, dtype=np.int32)


# ------------------------------------------------------------------------------------------------
# cameras and rays
def perspective(fovy_deg: float, aspect: float, near: float, far: float) -> np.ndarray:
    """glm::perspective (RH, -1..1 depth), as Player.cpp:6 uses it (fov 90, near 0.02, far 850)."""
    t = np.tan(np.deg2rad(fovy_deg) / 2.0)
    m = np.zeros((4, 4), dtype=np.float64)
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = -(far + near) / (far - near)
    m[3, 2] = -1.0
    m[2, 3] = -(2.0 * far * near) / (far - near)
    return m.astype(F32)


def look_at(eye, target, up=(0, 1, 0)) -> np.ndarray:
    eye, target, up = (np.asarray(a, dtype=np.float64) for a in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m.astype(F32)


def camera(eye, target, width: int, height: int, fovy_deg: float = 90.0):
    """Returns (inv_view, inv_proj) as float32 4x4 (row, column) matrices."""
    view = look_at(eye, target)
    proj = perspective(fovy_deg, width / height, 0.02, 850.0)
    return np.linalg.inv(view.astype(np.float64)).astype(F32), np.linalg.inv(proj.astype(np.float64)).astype(F32)


S260K_CAMERA = dict(eye=(-18.0, 5.0, 0.7), target=(10.0, 3.0, -0.4))


def random_rays(lo, hi, n: int, seed: int) -> np.ndarray:
    """Origins uniform in the box [lo, hi], directions uniform on the sphere (BASELINE config 5)."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, dtype=F32), np.asarray(hi, dtype=F32)
    o = (lo + (hi - lo) * rng.random((n, 3), dtype=F32)).astype(F32)
    z = (F32(2.0) * rng.random(n, dtype=F32) - F32(1.0)).astype(F32)
    ph = (F32(2.0 * np.pi) * rng.random(n, dtype=F32)).astype(F32)
    r = np.sqrt(np.maximum(F32(0.0), F32(1.0) - z * z)).astype(F32)
    d = np.stack([r * np.cos(ph), r * np.sin(ph), z], 1).astype(F32)
    rays = np.zeros(n, dtype=RAY_DT)
    rays["o"], rays["d"], rays["tmax"] = o, d, 1.0e6
    return rays


def _normalize(v):
    n = np.sqrt(np.sum(v * v, axis=1, keepdims=True)).astype(F32)
    return (v / np.maximum(n, F32(1e-30))).astype(F32)


def cos_weighted_hemisphere(n: np.ndarray, xi: np.ndarray) -> np.ndarray:
    """CosWeightedHemisphere of Shaders/Include/Sampling.glsl:1-12, vectorised."""
    uu = _normalize(np.cross(n, np.array([0.0, 1.0, 1.0], dtype=F32)).astype(F32))
    vv = np.cross(uu, n).astype(F32)
    ra = np.sqrt(xi[:, 1]).astype(F32)
    ang = (F32(2.0 * 3.14159265359) * xi[:, 0]).astype(F32)
    rx, ry = (ra * np.cos(ang)).astype(F32), (ra * np.sin(ang)).astype(F32)
    rz = np.sqrt(F32(1.0) - xi[:, 1]).astype(F32)
    return _normalize((rx[:, None] * uu + ry[:, None] * vv + rz[:, None] * n).astype(F32))


def bounce_rays(rays: np.ndarray, hits: np.ndarray, tris: np.ndarray, verts: np.ndarray, seed: int, spp: int = 1,
                offset: float = 0.05, tmax: float = 1.0e6):
    """Diffuse-GI style secondary rays (DiffuseTrace.glsl:445-446,:516-517): from every ray that hit,
    origin = P + N*offset and direction = CosWeightedHemisphere(N, xi), `spp` samples each.  N is the
    geometric normal turned against the incoming ray.  Rays that missed are dropped.
    Returns (new_rays, index of the parent ray for each new ray)."""
    ok = np.nonzero(hits["t"] > 0)[0]
    t = hits["t"][ok]
    P = (rays["o"][ok] + rays["d"][ok] * t[:, None]).astype(F32)
    tv = tris["v"][hits["tri"][ok]]
    A, B, Cc = (verts["position"][tv[:, k], :3] for k in range(3))
    N = _normalize(np.cross(B - A, Cc - A).astype(F32))
    flip = np.sum(N * rays["d"][ok], axis=1) > 0
    N[flip] = -N[flip]
    if spp > 1:
        P, N, ok = np.repeat(P, spp, 0), np.repeat(N, spp, 0), np.repeat(ok, spp)
    xi = np.random.default_rng(seed).random((len(P), 2), dtype=F32)
    out = np.zeros(len(P), dtype=RAY_DT)
    out["o"] = (P + N * F32(offset)).astype(F32)
    out["d"] = cos_weighted_hemisphere(N, xi)
    out["tmax"] = tmax
    return out, ok


def tile_order(width: int, height: int, tile: int = 64) -> np.ndarray:
    """Pixel indices grouped by tile x tile screen tiles, tiles in row-major order (config 4 shards tiles
    round-robin over ranks). Returns (pixel_index[W*H], tile_id[W*H])."""
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    tid = (ys // tile) * ((width + tile - 1) // tile) + (xs // tile)
    order = np.argsort(tid.ravel(), kind="stable")
    return order.astype(np.int64), tid.ravel()[order]
