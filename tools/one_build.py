"""Two builds of one scene (the second is the warm one), for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file out.csv python tools/one_build.py [--grid 2236]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from candela_b200 import api as cb, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=0, help="heightfield of 2 * grid^2 triangles instead of the 262k scene")
ap.add_argument("--fmt", type=int, default=0)
args = ap.parse_args()
v, i, m = scenes.make_heightfield(args.grid) if args.grid else scenes.make_s260k()
for rep in range(2):
    ri = cb.RayIntersector(args.fmt)
    ri.AddObject(2, v, i, m)
    print("triangles", len(i) // 3, "build ms", ri.last_build_ms)
    ri.close()
