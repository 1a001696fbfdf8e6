"""Writes profiles/r2_sass_excerpt.txt: the memory instructions of the default traversal kernels in the in-tree library
(cuobjdump -sass of candela_b200/libcandela_b200.so), so that the claims in DESIGN.md §4 stay checkable:
one LDG.E.ENL2.256 per node visit, no local-memory traffic in the stackless kernel, sm_100a only."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "candela_b200" / "libcandela_b200.so"
KERNELS = [("default stackless kernel, closest hit ignoring translucent entities (bench.py `value`)", r"trace_ww_stackless_kernelILi1ELi8ELi2ELb0ELb1E"),
           ("default stackless kernel, any hit (RTAO / shadow batches)", r"trace_ww_stackless_kernelILi2ELi8ELi2ELb0ELb1E"),
           ("default stack-format kernel, closest hit", r"trace_ww_stack_kernelILi0ELi2ELb0E"),
           ("staged kernel (top of the tree in shared memory), closest hit, 1024-thread CTAs", r"trace_hot_stackless_kernelILi0ELi1024ELi2E")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
    out = [f"# cuobjdump -sass candela_b200/libcandela_b200.so  (architectures in the fat binary: {', '.join(arch)})", ""]
    blocks = re.split(r"\n\s*Function : ", sass)
    for title, pat in KERNELS:
        blk = next((b for b in blocks if re.search(pat, b.split("\n", 1)[0])), None)
        if blk is None:
            out.append(f"## {title}: not found ({pat})")
            continue
        name = blk.split("\n", 1)[0].strip()
        ops = re.findall(r"\b(LDG[\w.]*|STG[\w.]*|LDL[\w.]*|STL[\w.]*|LDS[\w.]*|STS[\w.]*|ATOMG[\w.]*|RED[\w.]*|UTMALDG[\w.]*|UTCHMMA[\w.]*)\b", blk)
        counts = {}
        for o in ops:
            counts[o] = counts.get(o, 0) + 1
        m = re.search(re.escape(name) + r":\n\s*(REG:\d+ STACK:\d+ SHARED:\d+ LOCAL:\d+)", res)
        n_instr = len(re.findall(r"/\*[0-9a-f]{4,}\*/\s+\S", blk))
        out.append(f"## {title}")
        out.append(f"#  {name}")
        out.append(f"#  {m.group(1) if m else ''}   instructions: {n_instr}")
        for o, c in sorted(counts.items()):
            out.append(f"{c:6d}  {o}")
        first = [ln.strip() for ln in blk.splitlines() if "LDG.E.ENL2.256" in ln][:3]
        out += ["   e.g. " + f for f in first] + [""]
    (ROOT / "profiles" / "r2_sass_excerpt.txt").write_text("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
