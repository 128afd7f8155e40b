"""SE(2)-weighted nearest neighbour (Dubins state space) at bench scale: python profiles/se2_probe.py [ppc ...]"""
import sys, json
import numpy as np
import torch
sys.path.insert(0, ".")
from pyoptimalmotionplanning_b200 import MetricSpec
from pyoptimalmotionplanning_b200.backend import CudaNN
n, nq = 1 << 24, 1 << 20
pts = torch.from_numpy(np.random.default_rng(1234).random((n, 3)) * [1, 1, 2 * np.pi]).cuda()
q = torch.from_numpy(np.random.default_rng(4321).random((nq, 3)) * [1, 1, 2 * np.pi]).cuda()
st = torch.cuda.current_stream().cuda_stream
for ppc in [float(a) for a in sys.argv[1:]] or [0.0]:
    for k in (1, 16):
        nn = CudaNN(MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi]), 0)
        if ppc > 0:
            nn.set_param("points_per_cell", ppc)
        nn.set_device(pts.data_ptr(), n, st)
        nn.build_index(st)
        idx = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        dst = torch.empty((nq, k), dtype=torch.float64, device="cuda")
        nn.knearest_device(q.data_ptr(), nq, k, idx.data_ptr(), dst.data_ptr(), None, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nn.knearest_device(q.data_ptr(), nq, k, idx.data_ptr(), dst.data_ptr(), None, st)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"ppc": ppc, "k": k, "ms": e0.elapsed_time(e1), "cells": nn.get_stat("n_cells"),
                          "mean_dist": float(dst[:, 0].mean())}), flush=True)
        del nn
