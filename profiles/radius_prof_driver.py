"""Driver for launch lists / ncu captures of the radius query and the EST density (2^20 queries vs 2^24 points, ~16 hits)."""
import sys, math
import numpy as np
import torch
sys.path.insert(0, ".")
from pyoptimalmotionplanning_b200 import MetricSpec
from pyoptimalmotionplanning_b200.backend import CudaNN
n, nq = 1 << 24, 1 << 20
pts = torch.from_numpy(np.random.default_rng(1234).random((n, 2))).cuda()
q = torch.from_numpy(np.random.default_rng(4321).random((nq, 2))).cuda()
st = torch.cuda.current_stream().cuda_stream
nn = CudaNN(MetricSpec.euclidean(2), 0)
nn.set_device(pts.data_ptr(), n, st)
nn.build_index(st)
r = math.sqrt(16.0 / (n * math.pi))
off = torch.empty(nq + 1, dtype=torch.int64, device="cuda")
hits = torch.empty(nq * 40, dtype=torch.int64, device="cuda")
dens = torch.empty(nq, dtype=torch.float64, device="cuda")
for _ in range(2):
    nn.neighbors_device(q.data_ptr(), nq, r, True, off.data_ptr(), hits.data_ptr(), hits.numel(), st)
    nn.density_device(q.data_ptr(), nq, r, r / 4, True, dens.data_ptr(), None, st)
torch.cuda.synchronize()
print("hits", int(off[-1]))
