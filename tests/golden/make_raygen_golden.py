"""Generates tests/golden/raygen_golden.npz: OUTPUTS OF THE REFERENCE'S SHADER FUNCTIONS for the ray generators —
CosWeightedHemisphere, SampleGGXVNDF, SampleCone (Shaders/Include/Sampling.glsl), StochasticReflectionDirection
(Shaders/SpecularTrace.glsl:102-135) and ImportanceSample / LambertBRDF (Shaders/UpdateRadianceProbes.glsl:351-406) —
rewritten syntactically by oracle/ref_shim/glsl_to_cpp.py and compiled against the reference's glm (oracle/_ref, namespace
ref_raygen: hash2() bound to the counter stream, sin / cos / acos / pow bound to exact_math_ref.h).  The oracle's
restatement must return the same bits here; the fixture carries the compiled shaders' outputs to the GPU box.
Run only where /root/reference exists:   python tests/golden/make_raygen_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import binding as ob  # noqa: E402

PROBE_SEED = 77
PROBE_RES = (48, 24, 48)   # PROBE_GRID_X/Y/Z, Source/Core/Macros.h:22-24


def unit(rng, n):
    v = rng.normal(size=(n, 3)).astype(np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def main():
    ob.build_library(force=True)
    rng = np.random.default_rng(20261017)
    n = 6000
    N = unit(rng, n)
    N[:12] = [[0, 0, 1], [0, 1, 0], [1, 0, 0], [0, 0, -1], [0, -1, 0], [-1, 0, 0], [0, 0.70710678, 0.70710678], [0, -0.70710678, -0.70710678],
              [0.6, 0.8, 0], [0, 0.6, 0.8], [0.0004, 0.0003, 0.9999999], [0.001, 0.04, -0.9991994]]
    I = unit(rng, n)
    xi = rng.random((n, 2), dtype=np.float32)
    xi[:4] = [[0, 0], [0.99999994, 0.99999994], [0.5, 0], [0.25, 1.0]]
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    out = dict(normals=N, incident=I, xi=xi, keys=keys)
    cases = [("cos_hemisphere", ob.SAMPLE_COS_HEMISPHERE, dict(normals=N, xi=xi), 0.0),
             ("ggx_vndf_r0.27", ob.SAMPLE_GGX_VNDF, dict(normals=N, xi=xi), 0.27),
             ("ggx_vndf_r0.72", ob.SAMPLE_GGX_VNDF, dict(normals=N, xi=xi), 0.72),
             ("stochastic_reflection_r0.005", ob.SAMPLE_STOCHASTIC_REFLECTION, dict(normals=N, incident=I, keys=keys), 0.005),
             ("stochastic_reflection_r0.27", ob.SAMPLE_STOCHASTIC_REFLECTION, dict(normals=N, incident=I, keys=keys), 0.27),
             ("stochastic_reflection_r0.72", ob.SAMPLE_STOCHASTIC_REFLECTION, dict(normals=N, incident=I, keys=keys), 0.72),
             ("sample_cone_c0.98", ob.SAMPLE_CONE, dict(normals=N, xi=xi), 0.98),
             ("probe_importance_sample", ob.SAMPLE_PROBE, dict(keys=keys), 0.0)]
    total = 0
    for name, which, kw, rough in cases:
        ref = ob.ref_sample_directions(which, roughness=rough, **kw)
        mine = ob.sample_directions(which, roughness=rough, **kw)
        same = (ref.view(np.uint32) == mine.view(np.uint32)) | (np.isnan(ref) & np.isnan(mine))
        assert same.all(), name
        out[name] = ref
        total += len(ref)
    # the probe grid's directions: what cndl_generate_probe_rays_device must write, straight from the compiled ImportanceSample
    n_probe = PROBE_RES[0] * PROBE_RES[1] * PROBE_RES[2]
    pkeys = ob.stream_keys(PROBE_SEED, range(n_probe))
    pd = ob.ref_sample_directions(ob.SAMPLE_PROBE, keys=pkeys)
    assert pd.tobytes() == ob.sample_directions(ob.SAMPLE_PROBE, keys=pkeys).tobytes()
    out["probe_grid_directions"] = pd[::9].copy()   # every 9th probe keeps the file small
    total += len(pd)
    np.savez_compressed(ROOT / "tests" / "golden" / "raygen_golden.npz", **out)
    print("compiled-shader outputs for", total, "samples; oracle identical on every one")


if __name__ == "__main__":
    main()
