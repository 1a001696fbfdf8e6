"""Generates tests/golden/reference_traversal_golden.npz: OUTPUTS OF THE REFERENCE ITSELF — its GLSL traversal include
files (TraverseBVHStackless.glsl, TraverseBVHStack.glsl) rewritten syntactically by oracle/ref_shim/glsl_to_cpp.py and
compiled against its vendored glm (oracle/_ref/libcandela_ref.so) — for the cases of tests/reference_cases.py:
closest hit, closest hit ignoring translucent entities, any hit (1e6 and tmax 2.4), and GetData on the closest-hit records.
Run only where /root/reference exists:   python tests/golden/make_reference_traversal_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import binding as ob  # noqa: E402
import reference_cases  # noqa: E402


def main():
    ob.build_library(force=True)
    z = np.load(ROOT / "tests" / "golden" / "meshes.npz")
    names = sorted({k.split("__")[0] for k in z.files})
    meshes = {n: (z[f"{n}__p"].astype(np.float32), z[f"{n}__f"].astype(np.uint32)) for n in names}
    from candela_b200 import scenes
    meshes["dragon"] = scenes.load_dragon()
    out, total = {}, 0
    for fmt, label in ((ob.STACKLESS, "stackless"), (ob.STACK, "stack")):
        for c in reference_cases.build(ob, meshes, fmt):
            sc, rays = c["scene"], c["rays"]
            out[f"{label}/{c['name']}/rays"] = rays.view(np.float32).reshape(-1, 8)
            for kname, kind, tmax in reference_cases.KINDS:
                r = reference_cases.with_tmax(rays, tmax)
                ref = ob.ref_glsl_trace(fmt, kind, sc.nodes, sc.tris, sc.verts, sc.entities, r)
                mine, _ = sc.trace(kind, r)
                assert mine.tobytes() == ref.tobytes(), (label, c["name"], kname)
                out[f"{label}/{c['name']}/{kname}"] = ref.view(np.float32).reshape(len(r), -1)
                total += len(r)
            if fmt == ob.STACKLESS:
                hits = ob.ref_glsl_trace(fmt, 0, sc.nodes, sc.tris, sc.verts, sc.entities, rays)
                attr = ob.ref_glsl_get_data(sc.tris, sc.verts, sc.entities, hits)
                assert attr.tobytes() == ob.get_data(sc.tris, sc.verts, sc.entities, hits).tobytes(), (c["name"], "GetData")
                out[f"{label}/{c['name']}/get_data"] = attr.view(np.float32).reshape(len(hits), 8)
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_traversal_golden.npz", **out)
    print("reference outputs for", total, "ray queries; oracle identical on every one")


if __name__ == "__main__":
    main()
