"""The CUDA path AT THE BENCH CONFIGURATION (BASELINE configs[1]): 2^24 points, 2^20 queries,
D in {2,3,4,6}, k in {1,16}, the SE(2)-weighted metric and radius queries -- a sample of the answers of the
full-size batch is compared bit for bit with the oracle (the C restatement of pomp/structures/kdtree.py, which
is exact for Euclidean metrics, and the brute-force scan of nearestneighbors.py:88-141 for SE(2)).  This is
where int32 offsets, qi*k products, the overflow lists, the chunked host-buffer pipeline and its CUDA-graph
replay would break first; the small-N parity tests cannot see that.
"""
import math

import numpy as np
import pytest

from oracle import oracle
from pyoptimalmotionplanning_b200 import MetricSpec

pytestmark = pytest.mark.gpu

LOG2_N, LOG2_Q = 24, 20
SAMPLE = 4096


def _points(d, seed=1234):
    return np.random.default_rng(seed).random((1 << LOG2_N, d))


def _queries(d, seed=4321):
    return np.random.default_rng(seed).random((1 << LOG2_Q, d))


@pytest.fixture(scope="module")
def cache():
    return {}


def _setup(cache, d):
    """(points, queries, CudaNN with the index built, oracle kd-tree) per dimension, built once."""
    if d not in cache:
        from pyoptimalmotionplanning_b200.backend import CudaNN
        cache.clear()                       # one dimension resident at a time (host + device memory)
        pts, q = _points(d), _queries(d)
        nn = CudaNN(MetricSpec.euclidean(d), 0)
        nn.set(pts)
        kd = oracle.KDTree(MetricSpec.euclidean(d), pts)
        cache[d] = (pts, q, nn, kd)
    return cache[d]


@pytest.mark.parametrize("d", [2, 3, 4, 6])
def test_full_size_knn_sample_equals_oracle(cache, d):
    pts, q, nn, kd = _setup(cache, d)
    rng = np.random.default_rng(50 + d)
    # the sample always contains the first / last queries (chunk boundaries of the pipelined call)
    sub = np.unique(np.concatenate([rng.choice(1 << LOG2_Q, SAMPLE, replace=False), np.arange(64),
                                    np.arange((1 << LOG2_Q) - 64, 1 << LOG2_Q)]))
    for k in (1, 16):
        idx, dist, cnt = nn.knearest(q, k)                 # host buffers: the chunked, graph-replayed pipeline
        oi, od, oc = kd.knearest(q[sub], k)
        assert (cnt[sub] == oc).all()
        assert (idx[sub] == oi).all(), "k=%d: %d of %d sampled rows differ" % (k, int((idx[sub] != oi).any(1).sum()), len(sub))
        assert (dist[sub] == od).all()
        assert idx.min() >= 0 and idx.max() < (1 << LOG2_N)
        # size-independent properties on EVERY query: sorted by distance, distances re-derivable from the indices
        assert (np.diff(dist, axis=1) >= 0).all()
        chk = rng.choice(1 << LOG2_Q, 20000, replace=False)
        dd = pts[idx[chk, 0]] - q[chk]
        s = np.zeros(len(chk))
        for j in range(d):                                  # vectorops.distance: sequential, one rounding per op
            s = s + dd[:, j] * dd[:, j]
        assert (np.sqrt(s) == dist[chk, 0]).all()
        if k == 1:
            # the same batch a second and third time (graph capture, then replay) and through nearest_h
            for _ in range(2):
                i2, d2, _ = nn.knearest(q, 1)
                assert (i2 == idx).all() and (d2 == dist).all()
            i1, d1 = nn.nearest(q)
            assert (i1 == idx[:, 0]).all() and (d1 == dist[:, 0]).all()


def test_full_size_device_resident_call_equals_host_call(cache):
    """bench.py's timed call (pomp_nn_nearest on device buffers) against the host-buffer call checked above."""
    import torch
    pts, q, nn, kd = _setup(cache, 2)
    qd = torch.from_numpy(q).cuda()
    idx = torch.empty(1 << LOG2_Q, dtype=torch.int64, device="cuda")
    dst = torch.empty(1 << LOG2_Q, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    nn.nearest_device(qd.data_ptr(), 1 << LOG2_Q, idx.data_ptr(), dst.data_ptr(), st)
    torch.cuda.synchronize()
    sub = np.random.default_rng(9).choice(1 << LOG2_Q, SAMPLE, replace=False)
    oi, od = kd.nearest(q[sub])
    assert (idx.cpu().numpy()[sub] == oi).all() and (dst.cpu().numpy()[sub] == od).all()


def test_full_size_radius_sample_equals_oracle(cache):
    pts, q, nn, kd = _setup(cache, 2)
    r = math.sqrt(16.0 / ((1 << LOG2_N) * math.pi))         # ~16 hits per query
    off, hits = nn.neighbors(q, r)
    assert off[0] == 0 and off[-1] == len(hits) and (np.diff(off) >= 0).all()
    assert 14.0 < len(hits) / float(1 << LOG2_Q) < 18.0
    sub = np.random.default_rng(10).choice(1 << LOG2_Q, 1024, replace=False)
    for j in sub:
        want = np.sort(kd.neighbors(q[j], r))                # d <= radius (kdtree.py:347)
        got = hits[off[j]:off[j + 1]]
        assert len(got) == len(want) and (got == want).all()


def test_full_size_se2_sample_equals_bruteforce_oracle():
    from pyoptimalmotionplanning_b200.backend import CudaNN
    m = MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi])
    pts = _points(3) * [1, 1, 2 * np.pi]
    q = _queries(3) * [1, 1, 2 * np.pi]
    nn = CudaNN(m, 0)
    nn.set(pts)
    idx, dist = nn.nearest(q)
    sub = np.random.default_rng(11).choice(1 << LOG2_Q, 192, replace=False)
    oi, od = oracle.nn_nearest(m, pts, q[sub])
    # product metrics square through libm pow in the reference: distances agree to <= 2 ulp and an index may
    # differ only between candidates that close
    np.testing.assert_allclose(dist[sub], od, rtol=4e-16)
    bad = np.nonzero(idx[sub] != oi)[0]
    for b in bad:
        d_got = oracle.metric(m, pts[idx[sub][b]], q[sub][b])
        assert abs(d_got - od[b]) <= 4 * np.spacing(od[b])
    assert len(bad) <= 2
