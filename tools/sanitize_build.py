"""Small builds through every size class of the SAH builder, for compute-sanitizer (memcheck / racecheck):
compute-sanitizer --tool racecheck python tools/sanitize_build.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from candela_b200 import api as cb, scenes  # noqa: E402

v, i, m = scenes.make_heightfield(60)  # 7200 triangles
rng = np.random.default_rng(1)
P = rng.uniform(-1, 1, size=(3 * 3000, 3)).astype(np.float32)
soup_v = cb.make_vertices(P)
soup_i = np.arange(len(P), dtype=np.uint32)
for fmt in (cb.STACKLESS, cb.STACK):
    for split in (0, 64, 1000):
        for (vv, ii, mm) in ((v, i, m), (soup_v, soup_i, None)):
            ri = cb.RayIntersector(fmt)
            ri.set_tuning(9, 1)        # four tiny ranges per warp on every level (the default packs only large levels)
            ri.set_tuning(8, split)
            ri.AddObject(2, vv, ii, mm)
            n, t, _ = ri.read_buffers()
            print(fmt, split, len(n), len(t), flush=True)
            ri.close()
