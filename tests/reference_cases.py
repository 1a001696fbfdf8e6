"""The (scene, rays) cases of the committed reference-output fixture tests/golden/reference_traversal_golden.npz, shared by
the generator and the tests.  The scenes are the ones of cases.py (built by the oracle builder from committed meshes, itself
pinned byte-for-byte to the reference builder); rays are subsampled to keep the fixture small."""
import numpy as np

import cases

LIMIT = {"dragon_random": 4096, "dragon_axis_parallel": 1536, "multi_entity": 3072, "shared_edges": 4096}
KINDS = (("closest", 0, 0.0), ("ignore_transparent", 1, 0.0), ("any", 2, 0.0), ("any_tmax", 2, 2.4))


def build(ob, golden_meshes, fmt):
    out = []
    for c in cases.build_cases(ob, golden_meshes, fmt):
        rays = c["rays"]
        n = LIMIT.get(c["name"], 256)
        if len(rays) > n:
            rays = rays[:: max(1, len(rays) // n)][:n]
        out.append(dict(name=c["name"], scene=c["scene"], rays=np.ascontiguousarray(rays)))
    return out


def with_tmax(rays, tmax):
    r = rays.copy()
    r["tmax"] = tmax
    return r
