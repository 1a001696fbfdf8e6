import sys
sys.path.insert(0, "/root/repo")
from candela_b200 import api as cb, scenes
v, i, m = scenes.make_heightfield(2236)
for rep in range(2):
    ri = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(2, v, i, m)
    print("build ms", ri.last_build_ms)
    ri.close()
