"""Headline step (2^20 queries vs 2^24 points, 2-D, k = 1) timed WITHOUT the library's per-kernel profiling events:
python profiles/step_probe.py -- prints ms per step over 50 steps (CUDA events around the loop)."""
import sys
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
idx = torch.empty(nq, dtype=torch.int64, device="cuda")
dst = torch.empty(nq, dtype=torch.float64, device="cuda")
for prof in (0, 1, 0):
    nn.set_param("profile", prof)
    for _ in range(5):
        nn.nearest_device(q.data_ptr(), nq, idx.data_ptr(), dst.data_ptr(), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        nn.nearest_device(q.data_ptr(), nq, idx.data_ptr(), dst.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    print("profile events %d: %.4f ms per step" % (prof, e0.elapsed_time(e1) / 50))
