"""TEST INFRASTRUCTURE ONLY: the frame-level call (cndl_trace_frame) restated on the CPU with the oracle's pieces.

One diffuse-GI frame as the reference's DiffuseTrace.glsl runs it, with the G-buffer replaced by a camera-ray pass:
camera rays (Intersectors/TraverseBVHStack.glsl:414-421) -> IntersectRay -> `spp` cosine-hemisphere rays per hit pixel
from P + N*0.05 (:445-446) -> IntersectRayIgnoreTransparent (:484) -> per further bounce one ray per surviving path from
P + N*0.02 (:516-517) -> IntersectRay (:518).  Random stream element of (pixel, sample) = pixel * spp + sample under
seed + bounce, so a frame does not depend on how its pixels are sharded.  Every step is an oracle function (binding.py);
`tracer` lets bench.py's reference arm run the traversal steps through the compiled reference GLSL instead.
"""
from __future__ import annotations

import numpy as np

from . import binding as ob

HIT16_DT = np.dtype([("t", "<f4"), ("tri", "<i4"), ("v", "<f4"), ("w", "<f4")])
PIXEL_DT = np.dtype([("t", "<f4"), ("tri", "<i4"), ("v", "<f4"), ("w", "<f4"), ("ao", "<f4"), ("t_mean", "<f4"), ("rays", "<i4"), ("escaped", "<i4")])
OUT_HIT32, OUT_HIT16, OUT_PIXEL32 = 0, 1, 2


def _default_tracer(fmt, nodes, tris, verts, ents, threads):
    def trace(kind, rays):
        return ob.trace(fmt, kind, nodes, tris, verts, ents, rays, nthreads=threads)[0]
    return trace


def diffuse_batches(fmt, nodes, tris, verts, ents, inv_view, inv_proj, W, H, spp=1, bounces=1, seed=1, tracer=None, threads=None, pixels=None):
    """Yields (bounce, rays, ray ids, hits, kind) for every diffuse batch of the frame, plus the camera pass first as bounce -1.
    `pixels` (optional): only these pixel indices (e.g. one shard of the frame) are traced."""
    threads = threads or ob.hardware_threads()
    trace = tracer or _default_tracer(fmt, nodes, tris, verts, ents, threads)
    prim = ob.primary_rays(inv_view, inv_proj, W, H)
    pix_ids = np.arange(W * H, dtype=np.uint32)
    if pixels is not None:
        pix_ids = np.ascontiguousarray(pixels, dtype=np.uint32)
        prim = np.ascontiguousarray(prim[pix_ids])
    phits = trace(ob.CLOSEST, prim)
    yield -1, prim, pix_ids, phits, ob.CLOSEST
    src_rays, src_hits, src_ids = prim, phits, pix_ids
    for b in range(bounces):
        rays, _, rids = ob.generate_rays(src_rays, src_hits, tris, verts, ents, kind=ob.GEN_DIFFUSE, spp=spp if b == 0 else 1, seed=(seed + b) & 0xFFFFFFFF,
                                         offset=0.05 if b == 0 else 0.02, tmax=1.0e6, ids=src_ids)
        kind = ob.CLOSEST_IGNORE_TRANSPARENT if b == 0 else ob.CLOSEST
        hits = trace(kind, rays) if len(rays) else np.zeros(0, dtype=ob.HIT_DT)
        yield b, rays, rids, hits, kind
        src_rays, src_hits, src_ids = rays, hits, rids


def trace_frame(fmt, nodes, tris, verts, ents, inv_view, inv_proj, W, H, spp=1, bounces=1, seed=1, out_format=OUT_HIT16, tracer=None, threads=None,
                pixels=None):
    """Returns (records of the whole row-major frame, diffuse rays traced).  With `pixels`, only those pixels are traced; the rest of
    the frame keeps miss records / zeroed pixels (compare on `pixels` only)."""
    n_pix = W * H
    traced = 0
    if out_format in (OUT_HIT32, OUT_HIT16):
        assert bounces == 1
        out32 = np.zeros(n_pix * spp, dtype=ob.HIT_DT)
        for f in ("t", "u", "v", "w"):
            out32[f] = -1.0
        for f in ("mesh", "tri", "entity"):
            out32[f] = -1
    acc = None
    for b, rays, rids, hits, _ in diffuse_batches(fmt, nodes, tris, verts, ents, inv_view, inv_proj, W, H, spp, bounces, seed, tracer, threads, pixels):
        if b == -1:
            acc = np.zeros(n_pix, dtype=PIXEL_DT)
            acc["ao"] = 1.0
            acc["t_mean"] = -1.0
            for f in ("t", "tri", "v", "w"):
                acc[f][rids] = hits[f]
            continue
        traced += len(rays)
        if b == 0:
            if out_format in (OUT_HIT32, OUT_HIT16):
                out32[rids] = hits
            pix = (rids // spp).astype(np.int64)
            hit = hits["t"] > 0
            c = np.minimum(np.maximum(hits["t"] / np.float32(2.4), np.float32(0.0)), np.float32(1.0)).astype(np.float32)   # clamp(TUVW.x / 2.4f, 0, 1)
            term = np.where(hit, ob.xmath(3, c, np.full_like(c, np.float32(1.23))), np.float32(1.0)).astype(np.float32)  # DiffuseTrace.glsl:494
            ao_sum = np.zeros(n_pix, np.float32)
            t_sum = np.zeros(n_pix, np.float32)
            np.add.at(ao_sum, pix, term)                      # unbuffered: float32 additions in sample order
            np.add.at(t_sum, pix[hit], hits["t"][hit])
            n = np.bincount(pix, minlength=n_pix).astype(np.int32)
            n_hit = np.bincount(pix[hit], minlength=n_pix).astype(np.int32)
            with np.errstate(divide="ignore", invalid="ignore"):
                acc["ao"] = np.where(n > 0, ao_sum / n.astype(np.float32), np.float32(1.0))
                acc["t_mean"] = np.where(n_hit > 0, t_sum / n_hit.astype(np.float32), np.float32(-1.0))
            acc["rays"] = n
            acc["escaped"] = n - n_hit
        else:
            pix = (rids // spp).astype(np.int64)
            acc["rays"] += np.bincount(pix, minlength=n_pix).astype(np.int32)
            acc["escaped"] += np.bincount(pix[~(hits["t"] > 0)], minlength=n_pix).astype(np.int32)
    if out_format == OUT_HIT32:
        return out32, traced
    if out_format == OUT_HIT16:
        o = np.zeros(n_pix * spp, dtype=HIT16_DT)
        o["t"], o["tri"], o["v"], o["w"] = out32["t"], out32["tri"], out32["v"], out32["w"]
        return o, traced
    return acc, traced
