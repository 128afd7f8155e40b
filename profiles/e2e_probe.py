"""End-to-end nearest() with pinned host buffers (pomp_nn_nearest_h) over pipeline settings:
python profiles/e2e_probe.py -- ms per 2^20-query call for chunk counts / shapes / strip on-off."""
import sys, time, ctypes as C
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
    check(nn.lib, nn.lib.pomp_nn_nearest_h(nn._h, C.c_void_p(q_h.data_ptr()), nq, C.c_void_p(idx_h.data_ptr()),
                                           C.c_void_p(dst_h.data_ptr())))
for strip in (1, 0):
    nn.set_param("strip_search", strip)
    for shape in (1, 0):
        nn.set_param("pipeline_shape", shape)
        for chunks in (2, 3, 4, 5, 6, 8, 12):
            nn.set_param("pipeline_chunks", chunks)
            for _ in range(4):
                call()
            t0 = time.perf_counter()
            for _ in range(20):
                call()
            print("strip %d shape %d chunks %2d: %.4f ms" % (strip, shape, chunks, (time.perf_counter() - t0) / 20 * 1e3), flush=True)
