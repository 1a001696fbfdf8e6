"""Regenerates profiles/traffic.json (what bench.py reports as roofline.traffic / l1tex_sectors_per_ray / dram_frac_of_peak) and a
readable summary from ONE `ncu --set full` capture of the headline traversal kernel made with the CURRENT binary.

  on the GPU box:   python tools/ncu_traffic.py --capture          (runs bench.py under ncu for one launch; writes gpurun_out/)
  anywhere:         python tools/ncu_traffic.py --parse gpurun_out/r2_trace_ww_raw.csv --rays 2073558 --out profiles
"""
from __future__ import annotations

import argparse
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def parse(raw_csv: Path, rays: int, out_dir: Path, tag: str, source: str, write_traffic: bool = True):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    kernel = vals[col["Kernel Name"]] if "Kernel Name" in col else "?"
    got = {k: (vals[col[k]], units[col[k]]) for k in KEEP if k in col}
    dram = to_bytes(*got["dram__bytes_read.sum"]) + to_bytes(*got["dram__bytes_write.sum"])
    sectors = float(got["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"][0].replace(",", ""))
    wavefronts = float(got["l1tex__data_pipe_lsu_wavefronts.sum"][0].replace(",", "")) if "l1tex__data_pipe_lsu_wavefronts.sum" in got else None
    traffic = {"dram_bytes_per_launch": int(dram), "dram_bytes_per_ray": round(dram / rays, 1), "l1tex_sectors_per_ray": round(sectors / rays, 1),
               "l1tex_wavefronts_per_ray": None if wavefronts is None else round(wavefronts / rays, 1),
               "dram_throughput_pct": float(got["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]),
               "l1tex_throughput_pct": float(got["l1tex__throughput.avg.pct_of_peak_sustained_active"][0]),
               "lanes_per_instruction": float(got["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
               "kernel": kernel, "rays": rays, "source": source}
    if write_traffic:
        (out_dir / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    lines = [f"# {source}", f"# kernel: {kernel}", f"# rays per launch: {rays}"]
    lines += [f"{k:90s} {v[0]:>16s} {v[1]}" for k, v in got.items()]
    (out_dir / f"{tag}_ncu_full.txt").write_text("\n".join(lines) + "\n")
    print(json.dumps(traffic))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--capture", action="store_true")
    ap.add_argument("--parse", default="")
    ap.add_argument("--rays", type=int, default=0)
    ap.add_argument("--out", default=str(ROOT / "profiles"))
    ap.add_argument("--tag", default="r2_trace_ww")
    ap.add_argument("--kernel", default="trace_ww_stackless", help="--capture: kernel name regex")
    ap.add_argument("--skip", type=int, default=6, help="--capture: launches of that kernel to skip first")
    ap.add_argument("--cmd", default="", help="--capture: command to profile instead of the bench (quoted, run through the current interpreter)")
    ap.add_argument("--source", default="", help="--parse: description written into the summary")
    args = ap.parse_args()
    go = ROOT / "gpurun_out"
    if args.capture:
        go.mkdir(exist_ok=True)
        rep = go / f"{args.tag}.ncu-rep"
        target = args.cmd.split() if args.cmd else [str(ROOT / "bench.py"), "--steps", "1", "--warmup", "3", "--passes", "2", "--no-strong", "--build-reps", "0"]
        cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "--kernel-name", f"regex:{args.kernel}", "--launch-skip", str(args.skip),
               "--launch-count", "1", "-f", "-o", str(rep.with_suffix("")), sys.executable] + target
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        raw = go / f"{args.tag}_raw.csv"
        with open(raw, "w") as f:
            subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], check=True, stdout=f)
        print("captured", rep, raw)
        return
    parse(Path(args.parse), args.rays, Path(args.out), args.tag,
          args.source or "ncu --set full --clock-control none, one launch of the bench.py diffuse batch (tools/ncu_traffic.py --capture), B200",
          write_traffic=not args.source)


if __name__ == "__main__":
    main()
