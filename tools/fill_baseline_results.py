"""Rewrites BASELINE.md §5 ("Results") from the bench lines under profiles/ (rows that come from tools/run_configs.py
and tools/dev_bench.py runs are kept as text here; their sources are named in the table)."""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    p = ROOT / "BASELINE.md"
    s = p.read_text()
    n1 = json.loads((ROOT / "profiles" / "r1_bench_n1.json").read_text())
    n8 = json.loads((ROOT / "profiles" / "r1_bench_n8.json").read_text())
    new = f'''## 5. Results

Measured by this repo's harness on B200 (round 1; raw lines under `profiles/`, experiment history in
`profiles/r1_experiments.md`).  The reference publishes no number for this metric, so `vs_baseline` stays null.

| Config | GPU result | Roofline fraction (6550.1 GB/s measured, B_ray from the oracle's counters) | CPU beside it (same run, 16 host threads unless stated) | Parity |
|---|---|---|---|---|
| 2 — 1920×1080 diffuse rays, closest hit, stackless, 1 B200 (`profiles/r1_bench_n1.json`) | {n1["value"]:.0f} Mrays/s, {n1["ms_per_step"]:.4f} ms per batch; end to end from host buffers {n1["e2e"]["value"]:.0f} Mrays/s | {n1["roofline"]["frac"]:.3f} ({n1["roofline"]["bytes_per_ray"]} B/ray) | {n1["cpu_baseline"]["value"]:.1f} Mrays/s: the reference's GLSL traversal compiled against its glm | all {n1["parity"]["checked_rays"]:,} hit records bit-identical to the oracle and to the compiled reference shaders |
| 2 at 8 B200 (`profiles/r1_bench_n8.json`; 2 and 4 GPUs: `r1_bench_n2.json`, `r1_bench_n4.json`) | {n8["value"]:.0f} Mrays/s, {n8["ms_per_step"]:.4f} ms max over ranks | {n8["roofline"]["frac"]:.3f} per GPU | — | same check on every rank |
| 2, stack node format (`profiles/r1_experiments.md`) | 2441 Mrays/s | 0.89 (2382 B/ray) | — | bit-identical, including the 63-entry stack break |
| 3 — RTAO, 4 spp, tmax 2.4, any hit | 11.2 Grays/s | 1.47 (walk served from L1/L2) | — | all 8.29 M results bit-identical |
| 4 — 3840×2160, 8 spp, 4 bounces | 3.33 Grays/s on 1 GPU, 26.1 on 8 (`profiles/r1_configs_n{{1,8}}.jsonl`) | — | 1/16 subsample | 1/64 sample of every bounce bit-identical |
| 5 — 10 M triangles, 100 M random rays | 2.37 Grays/s on 1 GPU, 19.0 on 8 | 0.55 per GPU (scene beyond the L2) | 11.8 Mrays/s | 1 M-ray sample bit-identical |
| BVH build, 262,624 triangles | {n1["build"]["gpu_ms"]:.2f} ms (exact binned SAH, both formats; LBVH 0.31 ms) | latency-bound | reference builder 257 ms, oracle port {n1["cpu_baseline"]["build_ms_1thread"]:.0f} ms, one thread | node and triangle buffers byte-identical to the compiled reference builder |
| BVH build, 10 M triangles | 21.0 ms (LBVH 4.5 ms) | — | 13,005 ms | byte-identical |
'''
    p.write_text(s[: s.index("## 5. Results")] + new)


if __name__ == "__main__":
    main()
