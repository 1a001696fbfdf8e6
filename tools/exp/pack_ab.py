"""SAH build time with and without the packed tiny-range level step (CNDL_KNOB_BUILD_PACK_MIN): 262k, 2 M and 10 M triangles.
  python tools/exp/pack_ab.py"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import candela_b200 as cb
from candela_b200 import scenes

cases = [("s260k", scenes.make_s260k()), ("heightfield1000", scenes.make_heightfield(1000)), ("heightfield2236", scenes.make_heightfield(2236))]
for name, (v, i, m) in cases:
    for fmt in (cb.STACKLESS, cb.STACK):
        row = {"scene": name, "triangles": len(i) // 3, "fmt": fmt}
        for label, knob in (("never", 1 << 30), ("default", 0), ("always", 1)):
            ts = []
            for rep in range(4):
                ri = cb.RayIntersector(fmt)
                ri.set_tuning(9, knob)
                ri.AddObject(2, v, i, m)
                ts.append(ri.last_build_ms)
                ri.close()
            row[label] = round(min(ts[1:]), 3)
        print(json.dumps(row), flush=True)
