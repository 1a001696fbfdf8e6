"""Rewrites BASELINE.md §5 ("Results") from the round-2 lines under profiles/ (bench.py at 1 / 2 / 4 / 8 GPUs, both arms;
tools/run_configs.py; tools/trace_bench.py; tools/multi_bench.py)."""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
P = ROOT / "profiles"


def jl(name):
    return [json.loads(l) for l in (P / name).read_text().splitlines() if l.startswith("{")]


def main():
    p = ROOT / "BASELINE.md"
    s = p.read_text()
    b = {n: json.loads((P / f"r2_bench_n{n}.json").read_text()) for n in (1, 2, 4, 8)}
    r = {n: json.loads((P / f"r2_bench_ref_n{n}.json").read_text()) for n in (1, 8)}
    c1 = jl("r2_configs_n1.jsonl")
    c8 = jl("r2_configs_n8.jsonl")
    tb = jl("r2_trace_bench.jsonl")
    m8 = json.loads((P / "r2_multi_bench_n8.json").read_text())
    rtao = next(c for c in c1 if c["config"].startswith("rtao") and not c["octant_bucketed"])
    spec = next(c for c in c1 if c["config"].startswith("specular_rough"))
    shadow = next(c for c in c1 if c["config"].startswith("shadow"))
    soup1 = next(c for c in c1 if c["config"].startswith("heightfield") and c["rays_all_ranks"] == 100000000)
    soup8 = next(c for c in c8 if c["config"].startswith("heightfield"))
    stack = next(t for t in tb if t["fmt"] == "stack" and t["octant"] == 0)
    n1, n8 = b[1], b[8]
    ss = {n: b[n]["strong_scaling"] for n in b}
    g8 = json.loads((P / "r2_bench_gather_n8.json").read_text())["strong_scaling"]
    eff = {n: ss[1]["ms_per_step"] / (n * ss[n]["ms_per_step"]) for n in b}
    new = f'''## 5. Results

Measured by this repo's harness on B200 (round 2; raw lines under `profiles/r2_*`, experiment history in
`profiles/r2_experiments.md` and `r1_experiments.md`).  The reference publishes no number for this metric, so `vs_baseline` stays null.
"Fraction" = algorithmic bytes per second (B_ray from the oracle's counters) over the measured 6550.1 GB/s; on the L2-resident 262k scene that
is throughput served from L1/L2 (the measured limiter there is the L1 data pipe at 86 % of peak), on the 10 M-triangle scene it is a DRAM roofline.

| Config | GPU result | Fraction | CPU beside it (same box) | Parity |
|---|---|---|---|---|
| 2 — 1920×1080 diffuse rays, closest hit, stackless, 1 B200 (`r2_bench_n1.json`) | **{n1["value"]:.0f} Mrays/s**, {n1["details"]["ms_per_pass"]:.3f} ms per pass (batch as the generator emits it, octant-major); **end to end {n1["e2e"]["value"]:.0f} Mrays/s** through the frame-level call ({n1["e2e"]["ms_per_frame"]:.3f} ms per frame: camera pass + generator + traversal, 16 B per pixel to pinned host memory); rays staged through the host as in round 1: {n1["e2e_host_rays"]["value"]:.0f} | {n1["roofline"]["frac"]:.2f} ({n1["roofline"]["bytes_per_ray"]} B/ray); l1tex {n1["roofline"]["l1tex_pct_of_peak"]:.0f} % of peak, {n1["roofline"]["lanes_per_instruction"]} lanes per instruction | {r[1]["value"]:.1f} Mrays/s on {r[1]["cpu_baseline"]["cores"]} threads: the reference's GLSL traversal compiled against its glm (`r2_bench_ref_n1.json`) | the batch bit-identical to the oracle's generator; all {n1["parity"]["checked_rays"]:,} hit records bit-identical to the oracle and to the compiled reference shaders; the last end-to-end frame bit-identical to the CPU frame |
| 2 at 2 / 4 / 8 B200, one batch per GPU (`r2_bench_n{{2,4,8}}.json`) | {b[2]["value"]:.0f} / {b[4]["value"]:.0f} / **{n8["value"]:.0f} Mrays/s**; end to end {b[2]["e2e"]["value"]:.0f} / {b[4]["e2e"]["value"]:.0f} / {n8["e2e"]["value"]:.0f} (the host takes ~100 GB/s of device-to-host traffic in total) | {n8["roofline"]["frac"]:.2f} per GPU | {r[8]["value"]:.1f} Mrays/s on {r[8]["cpu_baseline"]["cores"]} threads, 8 batches per step | same checks on rank 0 |
| 2, stack node format (`r2_trace_bench.jsonl`) | {stack["mrays_s"]:.0f} Mrays/s ({stack["ms"]:.3f} ms) | {stack["frac"]:.2f} (2382 B/ray) | — | bit-identical, including the 63-entry stack break |
| 3 — RTAO, 4 spp, tmax 2.4, any hit (`r2_configs_n1.jsonl`) | {rtao["mrays_s"] / 1e3:.2f} Grays/s ({rtao["ms"]:.3f} ms for {rtao["rays"]:,} rays) | {rtao["roofline_frac"]:.2f} (walk served from L1/L2) | {rtao["cpu_mrays_s"]:.1f} Mrays/s (oracle port) | batch and all 8.29 M results bit-identical (also under `pytest -m gpu`) |
| specular (GGX, roughness 0.3) / shadow batches, 1080p | {spec["mrays_s"] / 1e3:.2f} / {shadow["mrays_s"] / 1e3:.2f} Grays/s | {spec["roofline_frac"]:.2f} / {shadow["roofline_frac"]:.2f} | — | bit-identical |
| 4 — 3840×2160, 8 spp, 4 bounces = {ss[1]["rays_per_step"] / 1e6:.1f} M rays, tiles dealt round-robin, **the final frame assembled on rank 0 inside the timed region** — every rank's resolve kernel stores its pixels into rank 0's frame over NVLink peer memory (CUDA IPC), then a one-element all-reduce (`strong_scaling` in `r2_bench_n*.json`; with the NCCL gather + untile instead: `r2_bench_gather_n*.json`) | {ss[1]["ms_per_step"]:.1f} / {ss[2]["ms_per_step"]:.1f} / {ss[4]["ms_per_step"]:.1f} / **{ss[8]["ms_per_step"]:.1f} ms** at 1 / 2 / 4 / 8 GPUs = {ss[1]["mrays_s"] / 1e3:.2f} … {ss[8]["mrays_s"] / 1e3:.1f} Grays/s; strong-scaling efficiency {eff[2]:.2f} / {eff[4]:.2f} / **{eff[8]:.2f}**; completion barrier {ss[8]["final_frame_gather_ms"]:.2f} ms at 8 GPUs (NCCL gather + untile of the 265 MB frame: {g8["final_frame_gather_ms"]:.2f} ms, {g8["ms_per_step"]:.1f} ms per frame) | — | — | the same pipeline at 480×270 bit-identical to the CPU frame; one shard in sixteen of the 4K frame under `pytest -m gpu` |
| 4 through one process driving all GPUs (`cndl_multi_*`, `r2_multi_bench_n8.json`) | {min(m8["frame_ms_staged_copy_pipelined"], m8["frame_ms_peer_stores_pipelined"]):.1f} ms per frame at 8 GPUs (two frames in flight) including the copy of the frame to the host (single GPU {m8["single_device_frame_ms"]:.1f}); scene replication device to device: {m8["replicate_2M_tris_ms"]:.2f} ms for 2 M triangles to 7 peers | — | — | frame bit-identical to the single-GPU frame |
| 5 — 10 M triangles, 100 M random rays | {soup1["mrays_s"] / 1e3:.2f} Grays/s on 1 GPU ({soup1["trace_ms_max_over_ranks"]:.1f} ms, in-call ray ordering counted), {soup8["mrays_s"] / 1e3:.1f} on 8 | **{soup1["roofline_frac_per_gpu"]:.2f}** per GPU (DRAM roofline, {soup1["bytes_per_ray"]} B/ray; round 1: 0.55) | {soup1["cpu_mrays_s"]:.1f} Mrays/s | 1 M-ray sample bit-identical |
| BVH build, 262,624 triangles | {n1["build"]["gpu_ms"]:.2f} ms (exact binned SAH, both formats; LBVH 0.31 ms) | per-level latency | reference builder {r[1]["cpu_baseline"]["build_ms_reference_builder_1thread"]:.0f} ms, oracle port {r[1]["cpu_baseline"]["build_ms_port_1thread"]:.0f} ms, one thread | node and triangle buffers byte-identical to the compiled reference builder |
| BVH build, 10 M triangles | **19.0 ms** (`tools/exp/pack_ab.py`, tiny ranges packed four to a warp; {soup8["gpu_build_ms"]:.1f} ms in the `r2_configs` runs made before that change; LBVH 3.6 ms) | — | {soup1["cpu_build_ms"]:,} ms | byte-identical |
'''
    p.write_text(s[: s.index("## 5. Results")] + new)


if __name__ == "__main__":
    main()
