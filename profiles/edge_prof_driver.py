import numpy as np, torch, sys
sys.path.insert(0, '.')
from pyoptimalmotionplanning_b200 import SpaceSpec
from pyoptimalmotionplanning_b200.backend import CudaSpace
rng = np.random.default_rng(8)
c = rng.random((9996, 2)); hs = 0.001 + rng.random((9996, 2)) * 0.003
w = 0.02
kink = [[0.3, 0, 0.5 + w * 0.5, 0.2 - w * 0.5], [0.5 + w * 0.5, 0, 0.7, 0.3 - w * 0.5], [0.3, 0.2 + w * 0.5, 0.5 - w * 0.5, 0.7], [0.5 - w * 0.5, 0.3 + w * 0.5, 0.7, 0.7]]
sp = SpaceSpec([0, 0], [1, 1], boxes=np.concatenate([np.array(kink), np.concatenate([c - hs, c + hs], 1)]))
space = CudaSpace(sp, 0)
ne = 1 << 22
a_h = np.random.default_rng(7).random((ne, 2))
dlt = np.random.default_rng(9).random((ne, 2)) - 0.5
b_h = np.clip(a_h + dlt * (0.15 / np.maximum(np.linalg.norm(dlt, axis=1, keepdims=True), 1e-12)), 0, 1)
a, b = torch.from_numpy(a_h).cuda(), torch.from_numpy(b_h).cuda()
ok = torch.empty(ne, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    space.edge_check_device(a.data_ptr(), b.data_ptr(), ne, 0.01, ok.data_ptr(), st)
torch.cuda.synchronize()
print("feasible frac", ok.float().mean().item())
