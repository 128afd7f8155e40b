"""Small driver for ncu captures of one k-NN configuration: python profiles/knn_prof_driver.py D K [log2N] [log2Q]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pyoptimalmotionplanning_b200 import MetricSpec
from pyoptimalmotionplanning_b200.backend import CudaNN
D, K = int(sys.argv[1]), int(sys.argv[2])
n = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 24)
nq = 1 << (int(sys.argv[4]) if len(sys.argv) > 4 else 20)
pts = torch.from_numpy(np.random.default_rng(1234).random((n, D))).cuda()
q = torch.from_numpy(np.random.default_rng(4321).random((nq, D))).cuda()
st = torch.cuda.current_stream().cuda_stream
nn = CudaNN(MetricSpec.euclidean(D), 0)
nn.set_device(pts.data_ptr(), n, st)
nn.build_index(st)
idx = torch.empty((nq, K), dtype=torch.int64, device="cuda")
dst = torch.empty((nq, K), dtype=torch.float64, device="cuda")
for _ in range(3):
    nn.knearest_device(q.data_ptr(), nq, K, idx.data_ptr(), dst.data_ptr(), None, st)
torch.cuda.synchronize()
print("ok", idx[0].tolist())
