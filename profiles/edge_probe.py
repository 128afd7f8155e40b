"""Edge-check timing on the Kink + 10 K boxes field (bench extras inputs): python profiles/edge_probe.py"""
import sys, json
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from pyoptimalmotionplanning_b200.backend import CudaSpace
space = CudaSpace(bench.kink10k_space_spec(), 0)
ne = 1 << 22
a_h, b_h = bench.rrt_edges(ne)
a, b = torch.from_numpy(a_h).cuda(), torch.from_numpy(b_h).cuda()
ok = torch.empty(ne, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for res in (0.01, 0.001):
    ms = bench.time_cuda(lambda: space.edge_check_device(a.data_ptr(), b.data_ptr(), ne, res, ok.data_ptr(), st), reps=5, warm=2)
    print(json.dumps({"res": res, "ms": ms, "edges_per_s": ne / ms * 1e3, "feasible": float(ok.float().mean())}))
