"""The scene and query boxes of the collide-query tests (shared by the fixture generator and the tests)."""
import numpy as np

from cases import scale_rot, translate


def build(ob, golden_meshes):
    """Stackless scene: dragon + soup, four entities (identity, translated, rotated+scaled, mirrored scale)."""
    P, F = golden_meshes["dragon"]
    Ps, Fs = golden_meshes["soup400"]
    sc = ob.Scene(ob.STACKLESS)
    sc.add_object(2, ob.make_vertices(P), F.ravel(), (np.arange(len(F)) % 3).astype(np.int32))
    sc.add_object(3, ob.make_vertices(Ps), Fs.ravel(), np.full(len(Fs), 7, np.int32))
    sc.push_entity(2)
    sc.push_entity(3, model=translate(2.5, 0.2, -0.4))
    sc.push_entity(2, model=scale_rot(0.6, 40.0, (-2.0, 0.3, 0.8)))
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[1, 1], m[2, 2] = -1.5, 0.8, 1.2
    m[:3, 3] = (0.5, 2.0, -1.0)
    sc.push_entity(3, model=m)
    return sc


def boxes(ob, n=4000, seed=5):
    rng = np.random.default_rng(seed)
    c = rng.uniform((-3.5, -1.0, -2.5), (4.0, 3.5, 2.5), size=(n, 3)).astype(np.float32)
    h = (10.0 ** rng.uniform(-2.6, -0.4, size=(n, 3))).astype(np.float32)     # half extents 0.0025 .. 0.4, anisotropic
    h[::7] = np.float32(0.01)                                                  # CollidePoint-sized boxes
    b = ob.make_boxes(c - h, c + h)
    b["min"][5::97] = b["max"][5::97]                                          # degenerate (point) boxes
    return b
