import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle binding (test infrastructure)."""
    from oracle import binding
    binding.build_library()
    return binding


@pytest.fixture(scope="session")
def golden_meshes():
    z = np.load(GOLDEN / "meshes.npz")
    names = sorted({k.split("__")[0] for k in z.files})
    out = {n: (z[f"{n}__p"].astype(np.float32), z[f"{n}__f"].astype(np.uint32)) for n in names}
    from candela_b200 import scenes
    out["dragon"] = scenes.load_dragon()
    return out


@pytest.fixture(scope="session")
def cb():
    """The product package; on the GPU box the CUDA library must be the thing that runs."""
    import candela_b200
    from candela_b200 import api
    api.load_library()
    return candela_b200


@pytest.fixture(scope="session")
def s260k(cb, ob):
    """The ~260k-triangle benchmark scene on the GPU (stackless) with the buffers the oracle needs."""
    from candela_b200 import scenes
    v, i, m = scenes.make_s260k()
    ri = candela = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(2, v, i, m)
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    nodes, tris, _ = ri.read_buffers()
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
    yield dict(ri=ri, v=v, i=i, m=m, nodes=nodes, tris=tris, ents=ents)
    candela.close()
