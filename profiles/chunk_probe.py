"""Device-resident nearest() on sub-batches of the 2^20 queries: how the search scales below the full batch
(the e2e pipeline searches chunk by chunk).  python profiles/chunk_probe.py"""
import sys, json
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
for m in (1 << 20, 1 << 19, 377487, 304000, 1 << 18, 220000, 147000, 1 << 16, 1 << 15):
    for _ in range(3):
        nn.nearest_device(q.data_ptr(), m, idx.data_ptr(), dst.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        nn.nearest_device(q.data_ptr(), m, idx.data_ptr(), dst.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"queries": m, "ms": e0.elapsed_time(e1) / 20}), flush=True)

# the same sub-batch searches while both copy engines are kept busy (as in the e2e pipeline)
h_in = torch.empty(4 << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(4 << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(4 << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(4 << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for m in (377487, 220000):
    torch.cuda.synchronize()
    for _ in range(40):
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        nn.nearest_device(q.data_ptr(), m, idx.data_ptr(), dst.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"queries": m, "ms_with_copies_running": e0.elapsed_time(e1) / 10}), flush=True)
