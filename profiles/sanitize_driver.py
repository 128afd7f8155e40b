"""Small pass over this round's kernels, written for `compute-sanitizer --tool memcheck` (closed on this pool: the
wrapper answers 86 without running it; the plain pass runs in ~2 s and exits 0).  Covers:
strip streaming kernel, warp-per-query k-NN (3-D / 6-D, k = 1 / 16, 8 / 16 / 32 lanes), radius, edge kernels with the
parked-point queue, the device roadmap, EST projection density."""
import sys
import numpy as np
sys.path.insert(0, ".")
import pyoptimalmotionplanning_b200 as P
from pyoptimalmotionplanning_b200 import MetricSpec, SpaceSpec
from pyoptimalmotionplanning_b200.backend import CudaNN, CudaSpace
from pyoptimalmotionplanning_b200.roadmap import DeviceRoadmap
rng = np.random.default_rng(0)
pts = rng.random((1 << 19, 2))
nn = CudaNN(MetricSpec.euclidean(2), 0)
nn.set(pts)
q = rng.random((1 << 16, 2))
nn.set_param("strip_min_per_tile", 0)
i1, d1 = nn.nearest(q)
print("strip launches", nn.get_stat("strip_launches"))
nn.knearest(q[:20000], 16)
nn.neighbors(q[:40000], 0.003)
for D in (3, 6):
    p = rng.random((60000, D))
    m = CudaNN(MetricSpec.euclidean(D), 0)
    m.set_param("grid_min_points", 1)
    m.set_param("warp_search", 2)
    m.set(p)
    for lanes in (0, 8, 16, 32):
        m.set_param("warp_lanes", lanes)
        for k in (1, 5, 16, 32):
            m.knearest(rng.random((3000, D)) * 1.2 - 0.1, k)
c = rng.random((800, 2)); hs = 0.004 + rng.random((800, 2)) * 0.01
sp = SpaceSpec([0, 0], [1, 1], boxes=np.concatenate([c - hs, c + hs], 1))
space = CudaSpace(sp, 0)
a = rng.random((60000, 2)); b = np.clip(a + (rng.random((60000, 2)) - 0.5) * 0.3, 0, 1)
for res in (0.01, 0.001):
    space.edge_check(a, b, res)
space.feasible(a)
small = CudaNN(MetricSpec.euclidean(2), 0)
small.set(pts[:50000])
rm = DeviceRoadmap(small, space)
print("roadmap edges", rm.build(6, 0.01))
dens = P.ProjectionHashDensity(4, ([0] * 4, [1] * 4))
dens.add_batch(rng.random((500, 4)))
dens.density_batch(rng.random((100, 4)))
print("sanitize driver done")
