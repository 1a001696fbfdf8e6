"""Per-kernel totals and the level-by-level list of one build from an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv): python tools/build_launch_summary.py gpurun_out/build_launches.csv"""
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        rows.append((re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", ""), v, row["Grid Size"]))
    idx = [i for i, r in enumerate(rows) if "tri_precompute" in r[0]]
    b = rows[idx[-1]:]
    print(f"last build: {sum(r[1] for r in b):.1f} us over {len(b)} launches (cold-cache, serialised times)")
    agg = {}
    for n, v, g in b:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v:9.1f} us {c:4d}x {n}")
    if "--levels" in sys.argv:
        for n, v, g in b:
            print(f"  {n:40s} {v:8.1f} us grid {g}")


if __name__ == "__main__":
    main(sys.argv[1])
