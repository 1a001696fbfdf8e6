"""Generates tests/golden/collide_golden.npz with the UNMODIFIED reference's Physics::CollideBox
(oracle/_ref/libcandela_ref.so, compiled from /root/reference/Source/Core/Physics.cpp by oracle/Makefile).
Run only where /root/reference exists:   python tests/golden/make_collide_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import binding as ob  # noqa: E402
import collide_scene  # noqa: E402


def main():
    ob.build_library(force=True)
    z = np.load(ROOT / "tests" / "golden" / "meshes.npz")
    from candela_b200 import scenes
    meshes = {"soup400": (z["soup400__p"].astype(np.float32), z["soup400__f"].astype(np.uint32)), "dragon": scenes.load_dragon()}
    sc = collide_scene.build(ob, meshes)
    b = collide_scene.boxes(ob)
    ref = ob.ref_collide_boxes(sc.nodes, sc.tris, sc.verts, sc.entities, b)
    mine = ob.collide_boxes(sc.nodes, sc.tris, sc.verts, sc.entities, b)
    assert np.array_equal(ref, mine["collided"]), "oracle restatement differs from the reference"
    print("boxes", len(b), "collided", int(ref.sum()))
    np.savez_compressed(ROOT / "tests" / "golden" / "collide_golden.npz", boxes=b.view(np.float32).reshape(-1, 8), collided=ref.astype(np.uint8),
                        oracle=mine.view(np.int32).reshape(-1, 4))


if __name__ == "__main__":
    main()
