"""Worker of tests/test_gpu_peer_frame.py: two processes (gloo) on ONE GPU; rank 0 owns the frame, both ranks trace their tiles into it
through the CUDA IPC mapping (sharding.PeerFrame), rank 0 compares with the frame traced by one context alone."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import candela_b200 as cb  # noqa: E402
from candela_b200 import api, scenes, sharding  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(0)
    P, F = scenes.load_dragon()
    v = np.zeros(len(P), dtype=api.VERTEX_DT)
    v["position"][:, :3] = P
    v["position"][:, 3] = 1.0
    ri = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(1, v, F.astype(np.uint32).ravel(), np.zeros(len(F), np.int32))
    ri.BufferData()
    ri.PushEntity(1)
    ri.BufferEntities()
    W, H = 200, 120
    lo, hi = P.min(0), P.max(0)
    eye = (lo + hi) / 2 + np.array([0.0, 0.1, 1.6], np.float32) * float(np.linalg.norm(hi - lo))
    iv, ip = scenes.camera(eye=tuple(eye), target=tuple((lo + hi) / 2), width=W, height=H)
    stream = torch.cuda.current_stream().cuda_stream
    for fmt, rec in ((api.FRAME_OUT_PIXEL32, 32), (api.FRAME_OUT_HIT16, 16)):
        kw = dict(spp=2, bounces=2, seed=3, tile=32, out_format=fmt) if fmt == api.FRAME_OUT_PIXEL32 else dict(spp=2, seed=3, tile=32, out_format=fmt)
        n_rec = ri.frame_records(cb.frame_params(iv, ip, W, H, **kw))
        peer = sharding.PeerFrame(ri, n_rec * rec, dst=0, device="cpu")
        if rank == 0:
            peer.tensor().fill_(0xEE)
        torch.cuda.synchronize()
        dist.barrier()
        ri.trace_frame_device(cb.frame_params(iv, ip, W, H, shard_index=rank, shard_count=world, **kw), peer.ptr, 0, stream)
        torch.cuda.synchronize()      # gloo does not order CUDA streams: the stores are complete before the barrier
        peer.complete()
        if rank == 0:
            got = peer.tensor().cpu().numpy().tobytes()
            want = ri.TraceFrame(cb.frame_params(iv, ip, W, H, **kw)).tobytes()
            print(f"PEER fmt={fmt} same={int(got == want)} bytes={len(got)}", flush=True)
        peer.close()
    ri.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
