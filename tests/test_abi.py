"""CPU tests of the drop-in boundary: the library builds, loads, exports every symbol the header
declares, and refuses to run without a GPU (there is no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from candela_b200 import _build, api
    _build.build()
    return api.load_library()


def declared_functions():
    text = (ROOT / "include" / "candela_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cndl_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/candela_b200.h but not exported"
    from candela_b200 import api
    assert sorted(api.EXPORTS) == names


def test_record_layouts_match_the_reference():
    from candela_b200 import api
    sizes = [d.itemsize for d in (api.VERTEX_DT, api.TRIANGLE_DT, api.NODE_DT, api.STACK_NODE_DT, api.ENTITY_DT, api.RAY_DT, api.HIT_DT)]
    assert sizes == [32, 16, 32, 64, 192, 32, 32]
    assert api.ENTITY_DT.fields["node_offset"][1] == 128 and api.ENTITY_DT.fields["data"][1] == 136


def test_abi_version(lib):
    assert lib.cndl_abi_version() == 2


def test_no_gpu_means_no_context(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.cndl_create(C.byref(h), 0, 0) == -3      # CNDL_ERR_NO_DEVICE: fails loudly, no fallback
    assert not h.value
    from candela_b200 import CandelaError, RayIntersector
    with pytest.raises(CandelaError):
        RayIntersector()
    # the stand-alone BuildBVH has no CPU path either
    import numpy as np
    from candela_b200 import BuildBVH, STACKLESS, make_vertices
    with pytest.raises(CandelaError) as e:
        BuildBVH(STACKLESS, make_vertices(np.eye(3, dtype=np.float32)), np.arange(3, dtype=np.uint32))
    assert e.value.code == -3


def test_bad_node_format_rejected(lib):
    from candela_b200 import CandelaError, RayIntersector
    with pytest.raises(CandelaError, match="can only be of type"):
        RayIntersector(node_format=7)


def test_product_does_not_import_the_oracle():
    for p in (ROOT / "candela_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"):
            assert "oracle" not in p.read_text().lower().replace("oracle-built", "").replace("the cpu oracle", ""), p


def test_scene_generators_are_deterministic():
    from candela_b200 import scenes
    a, b = scenes.make_s260k(), scenes.make_s260k()
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
    assert len(a[1]) // 3 == 262624
    r1, r2 = scenes.random_rays((-1, -1, -1), (1, 1, 1), 100, 5), scenes.random_rays((-1, -1, -1), (1, 1, 1), 100, 5)
    assert r1.tobytes() == r2.tobytes()
    v, i, m = scenes.make_heightfield(33)
    assert len(i) // 3 == 32 * 32 * 2 and i.max() == len(v) - 1
