"""Local search of the sharded path on one GPU: a slab [0,1] x [0, 1/G] holding 2^24 / G points, 1.04 * 2^20 queries
inside it (what every rank of a G-GPU run answers per step): python profiles/slab_probe.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pyoptimalmotionplanning_b200 import MetricSpec
from pyoptimalmotionplanning_b200.backend import CudaNN
st = torch.cuda.current_stream().cuda_stream
for G in (1, 2, 8):
    n, nq = (1 << 24) // G, int((1 << 20) * 1.04) + 1024
    rng = np.random.default_rng(G)
    pts = rng.random((n, 2)); pts[:, 1] /= G
    q = rng.random((nq, 2)); q[:, 1] /= G
    pts, q = torch.from_numpy(pts).cuda(), torch.from_numpy(q).cuda()
    nn = CudaNN(MetricSpec.euclidean(2), 0)
    nn.set_device(pts.data_ptr(), n, st)
    nn.build_index(st)
    idx = torch.empty(nq, dtype=torch.int64, device="cuda")
    dst = torch.empty(nq, dtype=torch.float64, device="cuda")
    nn.set_param("profile", 0)
    for _ in range(5):
        nn.nearest_device(q.data_ptr(), nq, idx.data_ptr(), dst.data_ptr(), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        nn.nearest_device(q.data_ptr(), nq, idx.data_ptr(), dst.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    nn.set_param("profile", 1)
    for _ in range(10):
        nn.nearest_device(q.data_ptr(), nq, idx.data_ptr(), dst.data_ptr(), st)
    torch.cuda.synchronize()
    print("G=%d: n=%d nq=%d  %.4f ms per call; search kernel %.4f ms; strip launches %d; cells %.0f; overflow %d" % (
        G, n, nq, e0.elapsed_time(e1) / 30, nn.get_stat("search_kernel_ms") / 10, nn.get_stat("strip_launches"),
        nn.get_stat("n_cells"), nn.get_stat("overflow_queries")))
    del nn, pts, q
