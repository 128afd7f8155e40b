"""Randomised parity fuzz of the grid k-NN / radius paths against the oracle (not collected by pytest):
    python tests/fuzz_gpu.py [seconds] [seed] [metrics]
Draws point sets (uniform, clustered, degenerate aspect ratios, duplicates), batch sizes on both sides of the
query-sort threshold, k in 1..32, masks, tombstones and unindexed tails, and compares every answer of a random
sample of the queries bit for bit with the brute-force oracle."""
import os
import sys
import time
import numpy as np

sys.path.insert(0, ".")
from oracle import oracle                                   # noqa: E402
from pyoptimalmotionplanning_b200 import MetricSpec        # noqa: E402
from pyoptimalmotionplanning_b200.backend import CudaNN    # noqa: E402


def draw_points(rng, n, D):
    kind = rng.integers(0, 5)
    if kind == 0:
        p = rng.random((n, D))
    elif kind == 1:      # clusters: dense cells next to empty ones
        c = rng.random((rng.integers(2, 40), D))
        p = c[rng.integers(0, len(c), n)] + rng.normal(0, 0.003 * (1 + 10 * rng.random()), (n, D))
    elif kind == 2:      # extreme aspect ratio
        p = rng.random((n, D)) * np.where(np.arange(D) == 0, 1000.0, 0.01)
    elif kind == 3:      # many exact duplicates / lattice points (ties)
        p = rng.integers(0, max(3, int(n ** (1.0 / D)) // 2), (n, D)).astype(np.float64) * 0.125
    else:                # a line (one coordinate constant)
        p = rng.random((n, D))
        p[:, -1] = 0.5
    return np.ascontiguousarray(p)


def canonical_knearest(pts, q, k, mask):
    """Brute force with the library-wide tie rule: k smallest by (distance, insertion index).  The distance is
    vectorops.distance (sequential sum of squares, then sqrt), bit for bit.  The reference's own KNearestResult
    has no canonical answer when several points tie at the k-th distance (it depends on eviction order), so the
    oracle port is used for the distances only."""
    nq, n = len(q), len(pts)
    idx = np.full((nq, k), -1, dtype=np.int64)
    dist = np.full((nq, k), np.inf)
    cnt = np.zeros(nq, dtype=np.int32)
    alive = np.ones(n, bool) if mask is None else (np.asarray(mask) == 0)
    ids = np.nonzero(alive)[0]
    P = pts[ids]
    for i in range(nq):
        s = np.zeros(len(P))
        for j in range(P.shape[1]):
            t = P[:, j] - q[i, j]
            s = s + t * t
        d = np.sqrt(s)
        kk = min(k, len(P))
        if kk == 0:
            continue
        part = np.argpartition(d, kk - 1)[:kk] if kk < len(P) else np.arange(len(P))
        thr = d[part].max()
        cand = np.nonzero(d <= thr)[0]
        order = cand[np.lexsort((ids[cand], d[cand]))][:kk]
        idx[i, :kk] = ids[order]
        dist[i, :kk] = d[order]
        cnt[i] = kk
    return idx, dist, cnt


def fuzz_metrics(budget, seed):
    """Generic metrics (SO(2) products with periodic grid axes, weighted, L1/Linf, cost lift) against the oracle.
    Product / weighted metrics square through libm pow in the reference: candidates within a few ulps of each other
    may swap, everything else must be identical."""
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    mk = {
        "se2": lambda: MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, float(rng.choice([0.5 / np.pi, 1.0, 0.01]))]),
        "pendulum": lambda: MetricSpec.geodesic_product([("SO2",), ("R", 1)], [1.0, float(rng.choice([1.0, 0.1]))]),
        "so2x2": lambda: MetricSpec.geodesic_product([("SO2",), ("SO2",), ("R", 1)], [1.0, 2.0, 1.0]),
        "weighted": lambda: MetricSpec.weighted_euclidean([1.0, 4.0, 0.25]),
        "l1": lambda: MetricSpec.l1(3),
        "linf": lambda: MetricSpec.linf(2),
        "cost_se2": lambda: MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi]).with_cost(1.0, 1.5),
    }
    rounds = checks = 0
    while time.time() < t_end:
        name = str(rng.choice(sorted(mk)))
        m = mk[name]()
        n = int(rng.choice([20000, 80000, 200000]))
        span = np.ones(m.dim)
        off = np.zeros(m.dim)
        if "se2" in name:
            span[2], off[2] = rng.choice([2 * np.pi, 8 * np.pi]), rng.choice([0.0, -4 * np.pi])
        if name == "pendulum":
            span[:], off[:] = [14, 20], [-7, -10]
        if name == "so2x2":
            span[:2] = 2 * np.pi
        if m.has_cost:
            span[-1] = 2.0
        pts = off + rng.random((n, m.dim)) * span
        nn = CudaNN(m, 0)
        nn.set_param("grid_min_points", 1)
        nn.set_param("grid_min_queries", 1)
        nn.set(pts)
        for _ in range(2):
            nq = int(rng.choice([100, 3000, 40000]))
            q = off + rng.random((nq, m.dim)) * span
            if "se2" in name:
                q[: nq // 5, 2] = rng.choice([0.0, 1e-9, -1e-9, 2 * np.pi, -2 * np.pi, 2 * np.pi - 1e-12], nq // 5)
            sub = rng.choice(nq, min(nq, 120), replace=False)
            k = int(rng.choice([1, 3, 8, 16]))
            idx, dist, cnt = nn.knearest(q, k)
            oi, od, oc = oracle.nn_knearest(m, pts, q[sub], k)
            exact = name in ("l1", "linf")
            ok = (cnt[sub] == oc).all()
            if exact:
                ok = ok and (idx[sub] == oi).all() and (dist[sub] == od).all()
            else:
                ok = ok and np.allclose(dist[sub], od, rtol=1e-14, atol=0)
                for r, c in zip(*np.nonzero(idx[sub] != oi)):
                    dg = oracle.metric(m, pts[idx[sub][r, c]], q[sub][r])
                    ok = ok and abs(dg - od[r, c]) <= 4 * np.spacing(od[r, c])
            if not ok:
                print("MISMATCH metrics", dict(seed=seed, round=rounds, metric=name, n=n, nq=nq, k=k))
                return 1
            checks += len(sub)
        rounds += 1
        del nn
    print("metric fuzz ok: %d rounds, %d queries checked (seed %d)" % (rounds, checks, seed))
    return 0


def main():
    if len(sys.argv) > 3 and sys.argv[3] == "metrics":
        return fuzz_metrics(float(sys.argv[1]), int(sys.argv[2]))
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    rounds = checks = 0
    while time.time() < t_end:
        D = int(rng.choice([2, 2, 2, 3, 4, 6]))
        n = int(rng.choice([20000, 60000, 200000, 500000]))
        pts = draw_points(rng, n, D)
        m = MetricSpec.euclidean(D)
        nn = CudaNN(m, 0)
        nn.set_param("grid_min_points", 1)
        nn.set_param("grid_min_queries", 1)
        if rng.random() < 0.5:
            nn.set_param("points_per_cell", float(rng.choice([0.5, 1.0, 1.5, 2.0, 3.0, 8.0])))
        n0 = n if rng.random() < 0.6 else n - int(rng.integers(1, 1500))   # the rest arrives after the index build
        nn.set(pts[:n0])
        nn.build_index()
        if n0 < n:
            nn.add(pts[n0:])
        mask = None
        if rng.random() < 0.4:
            mask = (rng.random(n) < rng.choice([0.01, 0.3, 0.9])).astype(np.uint8)
            nn.set_mask(mask)
        for _ in range(3):
            nq = int(rng.choice([70, 3000, 40000, 140000]))
            q = draw_points(rng, nq, D) if rng.random() < 0.3 else pts[rng.integers(0, n, nq)] + rng.normal(0, 1e-3, (nq, D))
            if rng.random() < 0.2:
                q[: nq // 4] = pts[rng.integers(0, n, nq // 4)]          # exact hits
            sub = rng.choice(nq, min(nq, 250), replace=False)
            k = int(rng.choice([1, 1, 1, 2, 4, 5, 8, 16, 32]))
            idx, dist, cnt = nn.knearest(q, k)
            oi, od, oc = canonical_knearest(pts, q[sub], k, mask)
            _, od2, oc2 = oracle.nn_knearest(m, pts, q[sub], k, mask)
            assert (od2 == od).all() and (oc2 == oc).all(), "numpy checker disagrees with the oracle on distances"
            ok = (cnt[sub] == oc).all() and (dist[sub] == od).all() and (idx[sub] == oi).all()
            if not ok:
                print("MISMATCH knearest", dict(seed=seed, round=rounds, D=D, n=n, n0=n0, nq=nq, k=k, masked=mask is not None))
                bad = np.nonzero((idx[sub] != oi).any(1) | (dist[sub] != od).any(1))[0][:3]
                for b in bad:
                    print(" query", q[sub][b], "got", idx[sub][b], dist[sub][b], "want", oi[b], od[b])
                return 1
            checks += len(sub)
            if rng.random() < (1.0 if os.environ.get("FUZZ_RADIUS_ALWAYS") else 0.3):
                r = float(np.median(od[:, -1][np.isfinite(od[:, -1])])) * 1.5 if np.isfinite(od[:, -1]).any() else 0.1
                qs = q[sub][:60]
                off, hits = nn.neighbors(qs, r, True)
                ooff, ohits = oracle.nn_neighbors(m, pts, qs, r, True, mask)
                if not (len(hits) == len(ohits) and (off == ooff).all() and (hits == ohits).all()):
                    print("MISMATCH neighbors", dict(seed=seed, round=rounds, D=D, n=n, n0=n0, r=r, masked=mask is not None,
                                                     ppc=nn.get_stat("n_indexed") / max(nn.get_stat("n_cells"), 1)))
                    for j in range(len(qs)):
                        g, w = hits[off[j]:off[j + 1]], ohits[ooff[j]:ooff[j + 1]]
                        if len(g) != len(w) or (g != w).any():
                            print(" query", j, qs[j], "got", len(g), g[:12], "want", len(w), w[:12])
                            extra = sorted(set(g.tolist()) ^ set(w.tolist()))[:6]
                            for e in extra:
                                print("   differs on", e, "dist", oracle.metric(m, pts[e], qs[j]), "r", r,
                                      "masked" if mask is not None and mask[e] else "")
                            break
                    return 1
        rounds += 1
        del nn
    print("fuzz ok: %d rounds, %d queries checked (seed %d)" % (rounds, checks, seed))
    return 0


if __name__ == "__main__":
    sys.exit(main())
