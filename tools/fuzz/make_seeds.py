"""Writes the seed files of tools/fuzz/fuzz_loader.cpp into a directory: two OBJ files (one with normals, UVs, groups and a
.mtl material library), and the glTF of tests/test_assets.py in its three containers, with materials and a texture."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))


def main(out: Path):
    import test_assets as ta
    out.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(3)
    P = rng.uniform(-1, 1, size=(30, 3)).astype(np.float32)
    F = rng.integers(0, 30, size=(40, 3))
    ta._write_obj(out / "tri.obj", P, F)
    N = rng.normal(size=(30, 3)).astype(np.float32)
    UV = rng.random((30, 2)).astype(np.float32)
    ta._write_obj(out / "mtl.obj", P, F, N=N, UV=UV, groups=[("red", 0, 15), ("tex", 15, 40)])
    text = (out / "mtl.obj").read_text()
    (out / "mtl.obj").write_text("mtllib mats.mtl\no thing\n" + text + "g poly\nf 1/1/1 2/2/2 3/3/3 4/4/4 5/5/5\nf -1 -2 -3\n")
    (out / "mats.mtl").write_text("newmtl red\nKd 0.8 0.1 0.1\nnewmtl tex\nKd 1 1 1\nmap_Kd albedo.png\nmap_Bump -bm 0.5 normal.png\n")
    for kind in ("bin", "uri", "glb"):
        ta._write_gltf(out, kind)
    print("seeds in", out)


if __name__ == "__main__":
    main(Path(sys.argv[1] if len(sys.argv) > 1 else "/tmp/cndl_fuzz"))
