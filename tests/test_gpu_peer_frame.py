"""CUDA IPC peer frame (cndl_ipc_*, sharding.PeerFrame): two processes store their tiles into one frame; it equals the unsharded frame."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
def test_two_processes_assemble_one_frame_through_peer_memory():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29641",
           str(ROOT / "tests" / "peer_frame_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("PEER ")]
    assert len(lines) == 2 and all(" same=1 " in l for l in lines), p.stdout[-2000:] + p.stderr[-2000:]
