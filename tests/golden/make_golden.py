"""Generates the committed fixtures.  Run only where /root/reference exists:

    python tests/golden/make_golden.py

1. Converts the reference's intact bundled meshes (OBJ: positions + triangulated faces only) into
   small .npz files: candela_b200/data/dragon_25k.npz and tests/golden/meshes.npz.
2. Runs the UNMODIFIED reference builder (oracle/_ref/libcandela_ref.so, compiled from
   /root/reference by oracle/Makefile) on every fixture mesh, checks that the oracle restatement
   is byte-identical, and records SHA-256 digests of the reference's buffers in
   tests/golden/builder_golden.json.  For the stackless format the reference flips children at
   random (std::random_device), so the check is: oracle + the reference's own flips == reference
   bytes; the digest stored is that of the oracle's unflipped buffer, which that check validates.
3. Stores small traversal vectors (rays + the oracle's hit records) in
   tests/golden/traversal_golden.npz.  These pin the ORACLE AGAINST ITSELF over time (and the GPU
   path against it on the GPU box); the reference's traversal is GLSL and cannot run here, so they
   are not reference outputs.
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import binding as ob  # noqa: E402

MODELS = Path("/root/reference/Source/Models")
HERE = Path(__file__).resolve().parent


def load_obj(path: Path):
    vs, fs = [], []
    for line in open(path, errors="replace"):
        if line.startswith("v "):
            vs.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("f "):
            p = [int(tok.split("/")[0]) for tok in line.split()[1:]]
            p = [i - 1 if i > 0 else len(vs) + i for i in p]
            for k in range(1, len(p) - 1):
                fs.append([p[0], p[k], p[k + 1]])
    return np.array(vs, np.float32), np.array(fs, np.uint32)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synthetic_meshes():
    """Edge-case meshes with >= 100 triangles (below that the reference itself crashes)."""
    rng = np.random.default_rng(11)
    out = {}
    # coplanar grid: zero extent on one axis, many equal centroids per bin
    n = 12
    g = np.stack(np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    pos = np.concatenate([g, np.zeros((len(g), 1), np.float32)], 1)
    a = (np.arange(n)[:, None] * (n + 1) + np.arange(n)[None, :]).ravel()
    out["coplanar_grid"] = (pos, np.concatenate([np.stack([a, a + n + 1, a + n + 2], 1), np.stack([a, a + n + 2, a + 1], 1)]).astype(np.uint32))
    # the same triangle 130 times: every split fails
    out["duplicates"] = (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.tile(np.array([[0, 1, 2]], np.uint32), (130, 1)))
    # random soup with a wide size distribution
    c = rng.uniform(-5, 5, size=(400, 1, 3))
    p = (c + rng.normal(0, 1, size=(400, 3, 3)) * rng.lognormal(-2, 1, size=(400, 1, 1))).reshape(-1, 3).astype(np.float32)
    out["soup400"] = (p, np.arange(1200, dtype=np.uint32).reshape(-1, 3))
    # points on a line: two degenerate axes
    t = np.arange(303, dtype=np.float32)
    out["collinear"] = (np.stack([t, np.zeros_like(t), np.zeros_like(t)], 1), np.arange(303, dtype=np.uint32).reshape(-1, 3))
    # mix of +0.0 and -0.0 coordinates (min/max argument order matters for the sign of zero)
    p = rng.uniform(-1, 1, size=(360, 3)).astype(np.float32)
    p[::7, 0] = 0.0
    p[3::7, 0] = -0.0
    p[::5, 1] = -0.0
    out["signed_zero"] = (p, np.arange(360, dtype=np.uint32).reshape(-1, 3))
    return out


def main():
    assert MODELS.exists(), "needs /root/reference"
    ob.build_library(force=True)
    assert ob.ref_sizes() == (32, 16, 32, 64)

    dragon = load_obj(MODELS / "dragon" / "dragon.obj")
    assert dragon[0].shape == (12500, 3) and dragon[1].shape == (25000, 3)
    (ROOT / "candela_b200" / "data").mkdir(exist_ok=True)
    np.savez_compressed(ROOT / "candela_b200" / "data" / "dragon_25k.npz", positions=dragon[0], faces=dragon[1].astype(np.uint16))

    meshes = {"dragon": dragon, "peach_castle": load_obj(MODELS / "marioc" / "Peach's Castle.obj"),
              "zelda_market": load_obj(MODELS / "zeldamarket" / "market place.obj")}
    meshes.update(synthetic_meshes())
    small = {k: v for k, v in meshes.items() if k != "dragon"}
    np.savez_compressed(HERE / "meshes.npz", **{f"{k}__p": v[0] for k, v in small.items()}, **{f"{k}__f": v[1] for k, v in small.items()})

    golden = {}
    for name, (P, F) in meshes.items():
        V = ob.make_vertices(P)
        mesh_no, t_off = 3, 5  # non-trivial GlobalMeshNumber and triangle offset
        entry = {"vertices": int(len(P)), "triangles": int(len(F))}
        for fmt, label in ((ob.STACK, "stack"), (ob.STACKLESS, "stackless")):
            rn, rt, rv = ob.ref_build(fmt, [(V, F, mesh_no)], t_offset=t_off)
            b = ob.build(fmt, V, F.ravel(), np.full(len(F), mesh_no, np.int32), t_offset=t_off,
                         adopt_flips_from=rn if fmt == ob.STACKLESS else None)
            assert rn.tobytes() == b.nodes.tobytes(), (name, label, "nodes differ")
            assert rt.tobytes() == b.tris.tobytes(), (name, label, "triangles differ")
            assert rv.tobytes() == V.tobytes(), (name, label, "vertices differ")
            if fmt == ob.STACKLESS:
                plain = ob.build(fmt, V, F.ravel(), np.full(len(F), mesh_no, np.int32), t_offset=t_off)
                entry[label] = {"nodes_sha256_unflipped": sha(plain.nodes), "tris_sha256": sha(rt), "n_nodes": int(len(rn)),
                                "reference_flips_in_this_run": int(b.flips)}
            else:
                entry[label] = {"nodes_sha256": sha(rn), "tris_sha256": sha(rt), "n_nodes": int(len(rn))}
            entry["stats"] = b.stats
        golden[name] = entry
        print(name, entry["triangles"], "tris ->", entry["stack"]["n_nodes"], "nodes: oracle == reference")

    # multi-mesh object: concatenation + mesh ids
    parts = [(ob.make_vertices(meshes["zelda_market"][0]), meshes["zelda_market"][1], 4),
             (ob.make_vertices(meshes["soup400"][0]), meshes["soup400"][1], 9)]
    cv, ci, cm = ob.concat_meshes(parts)
    rn, rt, rv = ob.ref_build(ob.STACK, parts, t_offset=0)
    b = ob.build(ob.STACK, cv, ci, cm)
    assert rn.tobytes() == b.nodes.tobytes() and rt.tobytes() == b.tris.tobytes() and rv.tobytes() == cv.tobytes()
    golden["multi_mesh(zelda_market+soup400)"] = {"stack": {"nodes_sha256": sha(rn), "tris_sha256": sha(rt), "n_nodes": int(len(rn))}}
    print("multi-mesh: oracle == reference")

    (HERE / "builder_golden.json").write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")

    # traversal vectors (oracle outputs; see module docstring)
    P, F = dragon
    V = ob.make_vertices(P)
    rng = np.random.default_rng(1234)
    lo, hi = P.min(0), P.max(0)
    n = 4096
    rays = np.zeros(n, dtype=ob.RAY_DT)
    rays["o"] = (lo + (hi - lo) * rng.random((n, 3), dtype=np.float32)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmax"] = 1.0e6
    out = {"rays": rays}
    for fmt, label in ((ob.STACKLESS, "stackless"), (ob.STACK, "stack")):
        sc = ob.Scene(fmt)
        sc.add_object(2, V, F.ravel(), np.zeros(len(F), np.int32))
        sc.push_entity(2)
        hits, c = sc.trace(ob.CLOSEST, rays)
        any_t, _ = sc.trace(ob.ANY, rays)
        out[f"{label}_hits"] = hits
        out[f"{label}_any"] = any_t
        out[f"{label}_counters"] = np.array([c["node_iters"], c["tri_tests"], c["capped"], c["hits"]], dtype=np.int64)
        print(label, c)
    np.savez_compressed(HERE / "traversal_golden.npz", **out)


if __name__ == "__main__":
    main()
