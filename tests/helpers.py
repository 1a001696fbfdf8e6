"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def single_entity_scene(ob, fmt, positions, faces, mesh_ids=None, model=None, **kw):
    sc = ob.Scene(fmt)
    sc.add_object(2, ob.make_vertices(positions), np.asarray(faces, np.uint32).ravel(),
                  np.zeros(len(faces), np.int32) if mesh_ids is None else mesh_ids, **kw)
    sc.push_entity(2, model=model)
    return sc


def rays_in_box(lo, hi, n, seed, tmax=1.0e6, dtype=None):
    from oracle.binding import RAY_DT
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    r = np.zeros(n, dtype=RAY_DT)
    r["o"] = (lo + (hi - lo) * rng.random((n, 3), dtype=np.float32)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    r["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    r["tmax"] = tmax
    return r


def assert_hits_equal(gpu, ref, rtol=1e-5):
    """The parity bar of BASELINE.json: indices bit-exact; t and barycentrics within 1e-5 relative.
    (The CUDA path is built to be bit-identical, so exact equality is checked first and reported.)"""
    for f in ("mesh", "tri", "entity", "iters"):
        bad = np.nonzero(gpu[f] != ref[f])[0]
        assert bad.size == 0, f"{f}: {bad.size} mismatches, first at ray {bad[:5]}: gpu {gpu[f][bad[:5]]} vs oracle {ref[f][bad[:5]]}"
    for f in ("t", "u", "v", "w"):
        a, b = gpu[f].astype(np.float64), ref[f].astype(np.float64)
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
        same_nan = np.isnan(a) & np.isnan(b)
        assert np.all((err <= rtol) | same_nan), f"{f}: max rel err {np.nanmax(err)}"


def bit_identical(gpu, ref) -> bool:
    return gpu.tobytes() == ref.tobytes()
