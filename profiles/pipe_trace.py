import sys, time, ctypes as C, os
import numpy as np
import torch
sys.path.insert(0, ".")
from pyoptimalmotionplanning_b200 import MetricSpec
from pyoptimalmotionplanning_b200.backend import CudaNN
from pyoptimalmotionplanning_b200._abi import check
n, nq = 1 << 24, 1 << 20
pts = torch.from_numpy(np.random.default_rng(1234).random((n, 2))).cuda()
q_h = torch.from_numpy(np.random.default_rng(4321).random((nq, 2))).pin_memory()
idx_h = torch.empty((nq, 1), dtype=torch.int64).pin_memory()
dst_h = torch.empty((nq, 1), dtype=torch.float64).pin_memory()
st = torch.cuda.current_stream().cuda_stream
nn = CudaNN(MetricSpec.euclidean(2), 0)
nn.set_device(pts.data_ptr(), n, st)
nn.build_index(st)
torch.cuda.synchronize()
def call():
    check(nn.lib, nn.lib.pomp_nn_nearest_h(nn._h, C.c_void_p(q_h.data_ptr()), nq, C.c_void_p(idx_h.data_ptr()), C.c_void_p(dst_h.data_ptr())))
for _ in range(6):
    call()
t0 = time.perf_counter()
for _ in range(20):
    call()
print("e2e %.4f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
os.environ["POMP_PIPE_TRACE"] = "1"
call()
