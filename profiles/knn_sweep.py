"""Timing sweep of the k-NN kernels: python profiles/knn_sweep.py [D,D,..] [k,k,..] -- per (D, k): thread-per-query
kernels (warp_search = 0) vs the warp-per-query shell kernel (warp_search = 2) over first-shell sizes and cell
occupancies.  2^20 queries vs 2^24 points, CUDA events, 3 warm-ups + 5 timed runs."""
import sys
import json
import numpy as np
import torch
sys.path.insert(0, ".")
from pyoptimalmotionplanning_b200 import MetricSpec
from pyoptimalmotionplanning_b200.backend import CudaNN

Ds = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 3, 4, 6]
Ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 16]
n, nq = 1 << 24, 1 << 20
st = torch.cuda.current_stream().cuda_stream
out = {}
for D in Ds:
    pts = torch.from_numpy(np.random.default_rng(1234).random((n, D))).cuda()
    q = torch.from_numpy(np.random.default_rng(4321).random((nq, D))).cuda()
    for K in Ks:
        idx = torch.empty((nq, K), dtype=torch.int64, device="cuda")
        dst = torch.empty((nq, K), dtype=torch.float64, device="cuda")
        ref = None
        variants = [("thread", 0, 0, 0)]
        for lanes in (0, 32):
            for ppc in (0,) + ((8,) if D >= 4 else ()):
                for need in (0,) + ((5,) if K == 1 else (1.6 * K,)):
                    variants.append(("warp%d" % lanes, ppc, need, 2))
        for name, ppc, need, ws in variants:
            nn = CudaNN(MetricSpec.euclidean(D), 0)
            nn.set_param("warp_search", ws)
            nn.set_param("warp_need", need)
            if name.startswith("warp"):
                nn.set_param("warp_lanes", int(name[4:]))
            if ppc:
                nn.set_param("points_per_cell", ppc)
            nn.set_device(pts.data_ptr(), n, st)
            nn.build_index(st)
            for _ in range(3):
                nn.knearest_device(q.data_ptr(), nq, K, idx.data_ptr(), dst.data_ptr(), None, st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                nn.knearest_device(q.data_ptr(), nq, K, idx.data_ptr(), dst.data_ptr(), None, st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            if ref is None:
                ref = (idx.clone(), dst.clone())
                same = True
            else:
                same = bool((idx == ref[0]).all()) and bool((dst == ref[1]).all())
            key = "D%d_k%d_%s_ppc%s_need%s" % (D, K, name, ppc, need)
            out[key] = {"ms": round(ms, 4), "same_as_thread_kernel": same}
            print(key, out[key], flush=True)
            del nn
    del pts, q
json.dump(out, open("gpurun_out/knn_sweep.json", "w"), indent=1)
