"""Driver for ncu captures of the generic-metric kernel (SE(2)-weighted nearest neighbour, 2^20 vs 2^24)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pyoptimalmotionplanning_b200 import MetricSpec
from pyoptimalmotionplanning_b200.backend import CudaNN
n, nq = 1 << 24, 1 << 20
pts = torch.from_numpy(np.random.default_rng(1234).random((n, 3)) * [1, 1, 2 * np.pi]).cuda()
q = torch.from_numpy(np.random.default_rng(4321).random((nq, 3)) * [1, 1, 2 * np.pi]).cuda()
st = torch.cuda.current_stream().cuda_stream
nn = CudaNN(MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi]), 0)
nn.set_device(pts.data_ptr(), n, st)
nn.build_index(st)
idx = torch.empty((nq, 1), dtype=torch.int64, device="cuda")
dst = torch.empty((nq, 1), dtype=torch.float64, device="cuda")
for _ in range(3):
    nn.knearest_device(q.data_ptr(), nq, 1, idx.data_ptr(), dst.data_ptr(), None, st)
torch.cuda.synchronize()
print("ok")
