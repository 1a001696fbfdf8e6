"""candela_b200 — B200-native (sm_100a) BVH build + ray traversal behind Candela's RayIntersector API.

The compute path is libcandela_b200.so (hand-written CUDA, C ABI in include/candela_b200.h).  This
package is the thin host-side mirror of the reference interface plus workload generators; it has no
CPU implementation of the path and raises if the library is missing.
"""
from . import api, scenes  # noqa: F401
from .api import (BUILDER_LBVH, BUILDER_SAH_EXACT, STACK, STACKLESS, SWAP_HASHED, SWAP_NONE, BuildBVH, CandelaError, PinnedBuffer,  # noqa: F401
                  MultiRayIntersector, RayIntersector, frame_params, make_rays, make_vertices)

__all__ = ["api", "scenes", "RayIntersector", "MultiRayIntersector", "frame_params", "BuildBVH", "CandelaError", "PinnedBuffer", "make_rays", "make_vertices", "STACKLESS", "STACK",
           "BUILDER_SAH_EXACT", "BUILDER_LBVH", "SWAP_NONE", "SWAP_HASHED"]
