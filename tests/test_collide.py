"""AABB collide query (Physics::CollideBox, Source/Core/Physics.cpp:21-228) — SURVEY.md §8f rank 4.
PINNED: the reference's Physics.cpp compiles here; the committed fixture holds its answers."""
import numpy as np
import pytest

import collide_scene

GOLDEN = "collide_golden.npz"


@pytest.fixture(scope="module")
def scene(ob, golden_meshes):
    return collide_scene.build(ob, golden_meshes)


def test_oracle_matches_reference_fixture(ob, scene):
    from conftest import GOLDEN as GDIR
    z = np.load(GDIR / GOLDEN)
    b = collide_scene.boxes(ob)
    assert b.view(np.float32).reshape(-1, 8).tobytes() == z["boxes"].tobytes(), "the fixture was made from other boxes"
    got = ob.collide_boxes(scene.nodes, scene.tris, scene.verts, scene.entities, b)
    assert np.array_equal(got["collided"], z["collided"].astype(np.int32))          # the reference's own answers
    assert got.view(np.int32).reshape(-1, 4).tobytes() == z["oracle"].tobytes()
    frac = got["collided"].mean()
    assert 0.05 < frac < 0.95 and len(np.unique(got["entity"][got["collided"] == 1])) == 4
    hit = got["collided"] == 1
    assert np.all(got["tri"][hit] >= 0) and np.all(got["mesh"][hit] == scene.tris["mesh"][got["tri"][hit]])
    assert np.all(got["tri"][~hit] == -1) and np.all(got["entity"][~hit] == -1)


def test_oracle_vs_reference_live(ob, scene):
    if not ob.REFERENCE_ROOT.exists():
        pytest.skip("/root/reference is not present (GPU box): the committed fixture covers this")
    b = collide_scene.boxes(ob, n=3000, seed=99)
    ref = ob.ref_collide_boxes(scene.nodes, scene.tris, scene.verts, scene.entities, b)
    assert np.array_equal(ref, ob.collide_boxes(scene.nodes, scene.tris, scene.verts, scene.entities, b)["collided"])


def test_known_answers(ob):
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5], [6, 5, 5], [5, 6, 5]], np.float32)
    F = np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
    sc = ob.Scene(ob.STACKLESS)
    sc.add_object(2, ob.make_vertices(P), F.ravel(), np.array([3, 4], np.int32))
    sc.push_entity(2)
    b = ob.make_boxes([(0.2, 0.2, -0.1), (2, 2, 2), (5.1, 5.1, 4.9), (0.2, 0.2, 0.5)], [(0.3, 0.3, 0.1), (3, 3, 3), (5.2, 5.2, 5.1), (0.3, 0.3, 0.6)])
    got = ob.collide_boxes(sc.nodes, sc.tris, sc.verts, sc.entities, b)
    assert list(got["collided"]) == [1, 0, 1, 0]
    assert got["mesh"][0] == 3 and got["mesh"][2] == 4 and got["entity"][0] == 0


@pytest.mark.gpu
def test_gpu_collide_bit_identical_to_oracle(cb, ob, scene):
    from conftest import GOLDEN as GDIR
    ri = cb.RayIntersector(cb.STACKLESS)
    off = 0
    for oid in (2, 3):
        o = scene.objects[oid]
        tris = scene.builds[oid].tris
        ri.AddPrebuiltObject(oid, scene.nodes[o["node_offset"]:o["node_offset"] + o["node_count"]], tris,
                             scene.verts[o["vert_offset"]:o["vert_offset"] + o["vert_count"]])
    ri.BufferData()
    ri.PushEntityRecords(scene.entities)
    ri.BufferEntities()
    b = collide_scene.boxes(ob)
    got = ri.CollideBoxes(b["min"], b["max"])
    want = ob.collide_boxes(scene.nodes, scene.tris, scene.verts, scene.entities, b)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(got["collided"], np.load(GDIR / GOLDEN)["collided"].astype(np.int32))   # == the reference
    k = int(np.nonzero(want["collided"] == 1)[0][0])
    assert ri.CollideBox(b["min"][k], b["max"][k]) is True
    c = 0.5 * (b["min"][k] + b["max"][k])
    assert ri.CollidePoint(c) == bool(ob.collide_boxes(scene.nodes, scene.tris, scene.verts, scene.entities, ob.make_boxes([c - np.float32(0.01)], [c + np.float32(0.01)]))["collided"][0])
    # device entry point
    import torch
    d_b = torch.from_numpy(b.view(np.float32).reshape(-1, 8)).cuda()
    d_o = torch.zeros((len(b), 4), dtype=torch.int32, device="cuda")
    ri.collide_boxes_device(d_b.data_ptr(), len(b), d_o.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert d_o.cpu().numpy().tobytes() == want.tobytes()
    st = cb.RayIntersector(cb.STACK)
    with pytest.raises(cb.CandelaError):
        st.CollideBoxes([(0, 0, 0)], [(1, 1, 1)])
    st.close()
    ri.close()
