import sys, numpy as np, torch
sys.path.insert(0, ".")
from pyoptimalmotionplanning_b200 import MetricSpec
from pyoptimalmotionplanning_b200.backend import CudaNN
n, nq, D = 1 << 24, 1 << 20, 6
pts = torch.from_numpy(np.random.default_rng(1234).random((n, D))).cuda()
q = torch.from_numpy(np.random.default_rng(4321).random((nq, D))).cuda()
st = torch.cuda.current_stream().cuda_stream
for ppc in (4, 8):
    nn = CudaNN(MetricSpec.euclidean(D), 0)
    nn.set_param("points_per_cell", ppc)
    nn.set_device(pts.data_ptr(), n, st)
    nn.build_index(st)
    for K in (1, 4, 8, 16):
        idx = torch.empty((nq, K), dtype=torch.int64, device="cuda")
        dst = torch.empty((nq, K), dtype=torch.float64, device="cuda")
        for _ in range(2):
            nn.knearest_device(q.data_ptr(), nq, K, idx.data_ptr(), dst.data_ptr(), None, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            nn.knearest_device(q.data_ptr(), nq, K, idx.data_ptr(), dst.data_ptr(), None, st)
        e1.record(); torch.cuda.synchronize()
        print("ppc", ppc, "k", K, "auto dispatch ms", round(e0.elapsed_time(e1) / 3, 3), flush=True)
