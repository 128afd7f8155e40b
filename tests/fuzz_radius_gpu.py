"""Targeted radius-query fuzz (hit counts on both sides of the 32-hit staging capacity):
    python tests/fuzz_radius_gpu.py [seconds] [seed]"""
import sys
import time
import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import oracle                                   # noqa: E402
from pyoptimalmotionplanning_b200 import MetricSpec        # noqa: E402
from pyoptimalmotionplanning_b200.backend import CudaNN    # noqa: E402
from fuzz_gpu import draw_points                            # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    rounds = 0
    while time.time() < t_end:
        D = int(rng.choice([2, 3, 4, 6]))
        n = int(rng.choice([30000, 200000, 500000]))
        pts = draw_points(rng, n, D)
        m = MetricSpec.euclidean(D)
        nn = CudaNN(m, 0)
        nn.set_param("grid_min_points", 1)
        nn.set_param("grid_min_queries", 1)
        n0 = n if rng.random() < 0.6 else n - int(rng.integers(1, 1500))
        nn.set(pts[:n0])
        nn.build_index()
        if n0 < n:
            nn.add(pts[n0:])
        mask = None
        if rng.random() < 0.3:
            mask = (rng.random(n) < 0.3).astype(np.uint8)
            nn.set_mask(mask)
        for _ in range(4):
            nq = int(rng.choice([60, 300, 5000]))
            q = pts[rng.integers(0, n, nq)] + rng.normal(0, 1e-3, (nq, D))
            k = int(rng.choice([8, 20, 32, 40, 64]))
            _, od, _ = oracle.nn_knearest(m, pts, q[:50], min(k, n), mask)
            fin = od[:, -1][np.isfinite(od[:, -1])]
            r = float(np.median(fin)) if len(fin) else 0.1
            inclusive = bool(rng.random() < 0.7)
            off, hits = nn.neighbors(q, r, inclusive)
            sub = rng.choice(nq, min(nq, 150), replace=False)
            ooff, ohits = oracle.nn_neighbors(m, pts, q[sub], r, inclusive, mask)
            for j, qi in enumerate(sub):
                got = hits[off[qi]:off[qi + 1]]
                want = ohits[ooff[j]:ooff[j + 1]]
                if len(got) != len(want) or (got != want).any():
                    print("MISMATCH", dict(seed=seed, round=rounds, D=D, n=n, n0=n0, nq=nq, r=r, inclusive=inclusive,
                                           masked=mask is not None, query=int(qi), got_len=len(got), want_len=len(want)))
                    print(" got ", got[:40])
                    print(" want", want[:40])
                    return 1
        rounds += 1
        del nn
    print("radius fuzz ok: %d rounds (seed %d)" % (rounds, seed))
    return 0


if __name__ == "__main__":
    sys.exit(main())
