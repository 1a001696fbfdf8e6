"""The C++17 host-side mirror (include/candela_b200/Intersector.hpp) compiles against the C ABI and,
on the GPU box, runs the reference's call sequence end to end."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "tests" / "cpp" / "mirror_demo"


def compile_demo():
    from candela_b200 import _build
    _build.build()
    cmd = ["g++", "-std=c++17", "-O1", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cpp" / "mirror_demo.cpp"), "-o", str(EXE),
           f"-L{ROOT / 'candela_b200'}", "-lcandela_b200", f"-Wl,-rpath,{ROOT / 'candela_b200'}"]
    subprocess.run(cmd, check=True, capture_output=True)


def test_mirror_compiles_and_fails_loudly_without_gpu():
    import torch
    compile_demo()
    p = subprocess.run([str(EXE)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert p.returncode == 0 and p.stdout.startswith("OK"), p.stdout + p.stderr
    else:
        assert p.returncode == 3 and p.stdout.startswith("NO_GPU"), p.stdout + p.stderr


@pytest.mark.gpu
def test_mirror_runs_the_reference_call_sequence():
    compile_demo()
    p = subprocess.run([str(EXE)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.strip().splitlines()
    assert lines[0].startswith("OK ")
    n_hit, total = int(lines[0].split()[1]), float(lines[0].split()[2])
    assert n_hit > 90 and abs(total - 3.0 * n_hit) < 1e-3      # floor at y = 0, rays from y = 3 straight down
    extra = [l for l in lines if l.startswith("EXTRA ")][0].split()
    assert extra[1] == f"data={n_hit}" and extra[2] == "collide=101"          # floor point collides, air point does not, box does
    mat = [l for l in lines if l.startswith("MATERIAL ")][0].split()
    # GenerateMeshTextureReferences: 6 albedo handles + 1 shared normal handle -> 7 indices, mesh 5 valid (index 6), mesh 0 not (-1)
    assert mat[1:] == [f"tex={n_hit}", "same=100", "handles=7", "ref5=6", "ref0=-1"]
    build = [l for l in lines if l.startswith("BUILDBVH ")][0].split()
    assert build[1] == "same=1" and build[3] == "tris=576"                    # BVH::BuildBVH free function: same bytes as AddObject's buffers
    frame = [l for l in lines if l.startswith("FRAME ")][0].split()
    assert int(frame[1].split("=")[1]) > 1000 and frame[2] in ("devices=2",) and frame[3] == "same=1"   # multi-device frame == single-device frame
    assert "THROW Trying to push entity whose parent object hasn't been added to global BVH" in p.stdout
