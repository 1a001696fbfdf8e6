"""Build times of the exact SAH builder for several multi-CTA split thresholds (knob 8), with a byte comparison
against the CPU oracle.  python tools/build_bench.py [--big] [--check]"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from candela_b200 import api as cb, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also the 2.0 M-triangle heightfield")
    ap.add_argument("--grid", type=int, default=0, help="also a heightfield of 2 * grid^2 triangles (2236 -> 10 M)")
    ap.add_argument("--fmts", default="0,1")
    ap.add_argument("--check", action="store_true", help="byte-compare with the CPU oracle (tests only)")
    ap.add_argument("--thresholds", default="0,2048,4096,8192,32768,65536,67108864")
    args = ap.parse_args()
    cases = [("s260k", scenes.make_s260k())]
    if args.big:
        cases.append(("heightfield1000", scenes.make_heightfield(1000)))
    if args.grid:
        cases.append((f"heightfield{args.grid}", scenes.make_heightfield(args.grid)))
    for name, (v, i, m) in cases:
        ref = None
        if args.check:
            from oracle import binding as ob
            ref = {f: ob.build(f, v, i, m) for f in (ob.STACKLESS, ob.STACK)}
        for fmt in [int(f) for f in args.fmts.split(',')]:
            for thr in [int(t) for t in args.thresholds.split(",")]:
                times = []
                ok = None
                for rep in range(5):
                    ri = cb.RayIntersector(fmt)
                    ri.set_tuning(8, thr)
                    ri.AddObject(2, v, i, m)
                    times.append(ri.last_build_ms)
                    if rep == 0 and ref is not None:
                        nodes, tris, _ = ri.read_buffers()
                        ok = nodes.tobytes() == ref[fmt].nodes.tobytes() and tris.tobytes() == ref[fmt].tris.tobytes()
                    ri.close()
                print(f"{name} fmt={fmt} split_node={thr}: build ms min {min(times):.3f} median {float(np.median(times)):.3f}"
                      + ("" if ok is None else f" byte-identical={ok}"), flush=True)


if __name__ == "__main__":
    main()
