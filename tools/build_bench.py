"""Developer tool: GPU build time on the S260k scene and (optionally) a heightfield, checked against the oracle."""
import argparse, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import candela_b200 as cb
from candela_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--heightfield", type=int, default=0, help="grid size n: (n-1)^2*2 triangles")
ap.add_argument("--check", action="store_true")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--builder", type=int, default=0)
args = ap.parse_args()
for name, (v, i, m) in (("s260k", scenes.make_s260k()),) + ((("heightfield%d" % args.heightfield, scenes.make_heightfield(args.heightfield)),) if args.heightfield else ()):
    for fmt in (cb.STACKLESS, cb.STACK):
        ms = []
        for _ in range(args.reps):
            ri = cb.RayIntersector(fmt)
            t0 = time.perf_counter()
            ri.AddObject(2, v, i, m, builder=args.builder)
            wall = 1e3 * (time.perf_counter() - t0)
            ms.append(ri.last_build_ms)
            if _ + 1 < args.reps:
                ri.close()
        line = f"{name} fmt={fmt} T={len(i)//3} nodes={ri.node_count} gpu_build_ms min={min(ms):.3f} med={sorted(ms)[len(ms)//2]:.3f} wall_ms_last={wall:.1f} launches={ri.launch_count}"
        if args.check and args.builder == 0:
            from oracle import binding as ob
            t0 = time.perf_counter()
            ref = ob.build(fmt, v, i, m)
            cpu = 1e3 * (time.perf_counter() - t0)
            nodes, tris, _v = ri.read_buffers()
            line += f" cpu_oracle_ms={cpu:.0f} identical={nodes.tobytes() == ref.nodes.tobytes() and tris.tobytes() == ref.tris.tobytes()}"
        print(line, flush=True)
        ri.close()
