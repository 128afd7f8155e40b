"""CUDA path vs the CPU oracle on seeded random inputs, through the C-ABI: bit-exact indices,
distances and booleans; edge cases the reference's two back ends exercise."""
import numpy as np
import pytest

from oracle import oracle
from pyoptimalmotionplanning_b200 import MetricSpec, SpaceSpec, DynamicsSpec, _abi

pytestmark = pytest.mark.gpu


def _nn(m, pts=None, grid=None):
    from pyoptimalmotionplanning_b200.backend import CudaNN
    nn = CudaNN(m, 0)
    if grid is True:
        nn.set_param("grid_min_points", 1)
        nn.set_param("grid_min_queries", 1)
    elif grid is False:
        nn.set_param("grid_min_points", 1e18)
    if pts is not None:
        nn.set(pts)
    return nn


@pytest.mark.parametrize("D", [2, 3, 4, 6])
@pytest.mark.parametrize("k", [1, 16])
def test_grid_knn_matches_oracle(D, k):
    rng = np.random.default_rng(100 + D)
    n, nq = 200000, 1500
    pts = rng.random((n, D))
    q = np.concatenate([rng.random((nq - 100, D)), rng.random((100, D)) * 1.4 - 0.2])  # some outside the bbox
    m = MetricSpec.euclidean(D)
    nn = _nn(m, pts, grid=True)
    idx, dist, cnt = nn.knearest(q, k)
    oi, od, oc = oracle.nn_knearest(m, pts, q, k)
    assert (cnt == oc).all()
    assert (idx == oi).all()
    assert (dist == od).all()
    # the brute-force kernels agree too
    nb = _nn(m, pts, grid=False)
    bi, bd, bc = nb.knearest(q[:200], k)
    assert (bi == oi[:200]).all() and (bd == od[:200]).all()


@pytest.mark.parametrize("k", [3, 33, 70, 128])
def test_k_buckets(k):
    rng = np.random.default_rng(7)
    pts = rng.random((5000, 3))
    q = rng.random((64, 3))
    m = MetricSpec.euclidean(3)
    oi, od, oc = oracle.nn_knearest(m, pts, q, k)
    for grid in (True, False):
        idx, dist, cnt = _nn(m, pts, grid).knearest(q, k)
        assert (idx == oi).all() and (dist == od).all() and (cnt == oc).all(), grid


def _metrics():
    return {
        "se2": MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi]),
        "pendulum": MetricSpec.geodesic_product([("SO2",), ("R", 1)], [1.0, 1.0]),
        "cv": MetricSpec.geodesic_product([("R", 2), ("R", 2)], [1.0, 1.0]),
        "l1": MetricSpec.l1(3),
        "linf": MetricSpec.linf(3),
        "weighted": MetricSpec.weighted_euclidean([1.0, 4.0, 0.25]),
        "scaled": MetricSpec.scaled_euclidean(2, 0.3),
        "cost_se2": MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi]).with_cost(1.0, 1.5),
        "cost_euclid": MetricSpec.euclidean(2).with_cost(0.7, None),
        "cost_off": MetricSpec.euclidean(3).with_cost(None, None),
    }


def _hits_equal_mod_boundary(m, pts, q, r, off, hits, ooff, ohits):
    """Radius hits as SETS per query; a point may be missing / extra only if its distance is within 2 ulp of the
    radius (x*x on the device vs libm pow(x, 2) in the reference)."""
    assert len(off) == len(ooff)
    for j in range(len(off) - 1):
        a, b = set(hits[off[j]:off[j + 1]].tolist()), set(ohits[ooff[j]:ooff[j + 1]].tolist())
        for i in a ^ b:
            d = oracle.metric(m, pts[i], q[j])
            if abs(d - r) > 2 * np.spacing(r):
                return False
        got = hits[off[j]:off[j + 1]]
        if (np.diff(got) <= 0).any():     # insertion order, no duplicates
            return False
    return True


def _near_tie_ok(m, pts, q, got_idx, want_idx, want_dist):
    """Metrics that square through libm pow in the reference can flip candidates that are within
    a couple of ulps of each other; accept those and nothing else."""
    bad = np.nonzero(got_idx != want_idx)
    for r, c in zip(*bad):
        d_got = oracle.metric(m, pts[got_idx[r, c]], q[r])
        if abs(d_got - want_dist[r, c]) > 4 * np.spacing(want_dist[r, c]):
            return False
    return len(bad[0]) <= max(2, got_idx.size // 500)


@pytest.mark.parametrize("name", sorted(_metrics()))
@pytest.mark.parametrize("grid", [True, False], ids=["grid", "brute"])
def test_general_metrics_match_oracle(name, grid):
    m = _metrics()[name]
    rng = np.random.default_rng(hash(name) % 1000)
    n, nq, k = 30000, 400, 5
    lo = np.zeros(m.dim)
    hi = np.ones(m.dim)
    if name in ("se2", "cost_se2"):
        hi[2] = 2 * np.pi
    if name == "pendulum":
        lo[:] = [-7, -10]
        hi[:] = [7, 10]
    if m.has_cost:
        hi[-1] = 2.0
    pts = lo + rng.random((n, m.dim)) * (hi - lo)
    q = lo + rng.random((nq, m.dim)) * (hi - lo)
    nn = _nn(m, pts, grid)
    idx, dist, cnt = nn.knearest(q, k)
    oi, od, oc = oracle.nn_knearest(m, pts, q, k)
    assert (cnt == oc).all()
    exact = m.kind not in (_abi.METRIC_WEIGHTED_EUCLIDEAN, _abi.METRIC_GEODESIC_PRODUCT)
    if exact:
        assert (idx == oi).all() and (dist == od).all()
    else:
        assert _near_tie_ok(m, pts, q, idx, oi, od)
        np.testing.assert_allclose(dist, od, rtol=1e-14)
    off, hits = nn.neighbors(q[:100], 0.05 * float(np.max(hi - lo)))
    ooff, ohits = oracle.nn_neighbors(m, pts, q[:100], 0.05 * float(np.max(hi - lo)), True)
    if exact:
        assert (off == ooff).all() and (hits == ohits).all()
    else:
        assert _hits_equal_mod_boundary(m, pts, q[:100], 0.05 * float(np.max(hi - lo)), off, hits, ooff, ohits)


def test_masks_tombstones_and_tail():
    rng = np.random.default_rng(3)
    D = 2
    m = MetricSpec.euclidean(D)
    pts = rng.random((60000, D))
    q = rng.random((800, D))
    nn = _nn(m, pts[:50000], grid=True)
    nn.build_index()
    nn.add(pts[50000:51500])          # unindexed tail (< max_tail)
    mask = np.zeros(51500, np.uint8)
    dead = rng.choice(51500, 5000, replace=False)
    for i in dead[:50]:
        nn.remove(int(i))
    mask[dead] = 1
    rej = np.zeros(51500, np.uint8)
    rej[dead[50:]] = 1
    nn.set_mask(rej)
    idx, dist, cnt = nn.knearest(q, 8)
    oi, od, oc = oracle.nn_knearest(m, pts[:51500], q, 8, mask)
    assert (idx == oi).all() and (dist == od).all()
    nn.add(pts[51500:])               # tail now exceeds max_tail -> rebuild on the next batch
    nn.set_mask(None)
    mask2 = np.zeros(60000, np.uint8)
    mask2[dead[:50]] = 1
    idx, dist = nn.nearest(q)
    oi, od = oracle.nn_nearest(m, pts, q, mask2)
    assert (idx == oi).all() and (dist == od).all()
    off, hits = nn.neighbors(q, 0.01)
    ooff, ohits = oracle.nn_neighbors(m, pts, q, 0.01, True, mask2)
    assert (off == ooff).all() and (hits == ohits).all()


def test_ties_duplicates_and_degenerate_sets():
    m = MetricSpec.euclidean(2)
    # lattice with duplicates: ties everywhere; lowest insertion index must win
    g = np.stack(np.meshgrid(np.arange(40) / 40.0, np.arange(40) / 40.0), -1).reshape(-1, 2)
    pts = np.concatenate([g, g[::3]])
    q = np.concatenate([g[::7] + 0.0125, g[::5]])
    # canonical rule under exact ties: the k smallest (distance, insertion index) keys.  (The
    # reference's KNearestResult evicts an arbitrary one of several tied maxima -- knn.py:20-23 --
    # and its kd-tree visits leaves in traversal order, so the reference itself has no single
    # answer here; the distances are still the oracle's, bit for bit.)
    d0 = pts[None, :, 0] - q[:, None, 0]
    d1 = pts[None, :, 1] - q[:, None, 1]
    dd = np.sqrt(d0 * d0 + d1 * d1)
    assert dd[3, 5] == oracle.metric(m, pts[5], q[3])
    order = np.lexsort((np.broadcast_to(np.arange(len(pts)), dd.shape), dd), axis=1)[:, :6]
    oi = order
    od = np.take_along_axis(dd, order, axis=1)
    for grid in (True, False):
        idx, dist, cnt = _nn(m, pts, grid).knearest(q, 6)
        assert (idx == oi).all() and (dist == od).all(), grid
        i1, dn = _nn(m, pts, grid).nearest(q)
        assert (i1 == oi[:, 0]).all()
        assert (i1 == oracle.nn_nearest(m, pts, q)[0]).all()
    # all points identical (zero extent): index has no usable axis -> brute-force kernels
    same = np.full((3000, 2), 0.25)
    nn = _nn(m, same, grid=True)
    idx, dist = nn.nearest(np.array([[0.3, 0.3]] * 70))
    assert (idx == 0).all()
    # empty set, k > n, single point
    nn = _nn(m)
    idx, dist = nn.nearest(np.zeros((3, 2)))
    assert (idx == -1).all() and np.isinf(dist).all()
    i1, d1 = nn.nearest_one([0.1, 0.2])
    assert i1 == -1
    nn.add(np.array([[0.5, 0.5], [0.25, 0.5]]))
    idx, dist, cnt = nn.knearest(np.array([[0.0, 0.5]]), 5)
    assert cnt.tolist() == [2] and idx[0].tolist() == [1, 0, -1, -1, -1]
    off, hits = nn.neighbors(np.zeros((0, 2)), 1.0)
    assert off.tolist() == [0] and len(hits) == 0


def test_radius_large_hit_counts_sorted():
    rng = np.random.default_rng(11)
    m = MetricSpec.euclidean(3)
    pts = rng.random((40000, 3))
    q = rng.random((128, 3))
    nn = _nn(m, pts, grid=True)
    off, hits = nn.neighbors(q, 0.2)
    ooff, ohits = oracle.nn_neighbors(m, pts, q, 0.2, True)
    assert (off == ooff).all() and (hits == ohits).all()


def test_edge_check_obstacle_grid_matches_oracle():
    rng = np.random.default_rng(8)
    c = rng.random((9996, 2))
    hs = 0.001 + rng.random((9996, 2)) * 0.003
    boxes = np.concatenate([c - hs, c + hs], axis=1)
    kink = np.array([[0.3, 0, 0.51, 0.19], [0.51, 0, 0.7, 0.29], [0.3, 0.21, 0.49, 0.7], [0.49, 0.31, 0.7, 0.7]])
    circles = np.concatenate([rng.random((300, 2)), 0.002 + rng.random((300, 1)) * 0.01], axis=1)
    sp = SpaceSpec([0, 0], [1, 1], boxes=np.concatenate([kink, boxes]), circles=circles)
    from pyoptimalmotionplanning_b200.backend import CudaSpace
    s = CudaSpace(sp, 0)
    x = rng.random((20000, 2)) * 1.1 - 0.05
    x[:50] = boxes[:50, :2]          # exactly on obstacle corners
    x[50:60, 0] = np.nan
    assert (s.feasible(x) == oracle.feasible(sp, x)).all()
    a = rng.random((20000, 2))
    d = rng.random((20000, 2)) - 0.5
    b = np.clip(a + d * (0.15 / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-9)) * rng.random((20000, 1)), 0, 1)
    for res in (0.01, 0.001):
        assert (s.edge_check(a, b, res) == oracle.edge_check(sp, a, b, res)).all()
    # indexed form over a node table
    ij = rng.integers(0, 20000, (30000, 2)).astype(np.int32)
    import torch
    nodes = torch.from_numpy(a).cuda()
    tij = torch.from_numpy(ij).cuda()
    ok = torch.empty(len(ij), dtype=torch.uint8, device="cuda")
    s.edge_check_indexed_device(nodes.data_ptr(), tij.data_ptr(), len(ij), 0.01, ok.data_ptr(),
                                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (ok.cpu().numpy().astype(bool) == oracle.edge_check_indexed(sp, a, ij, 0.01)).all()


def test_higher_dim_space_and_path_edges():
    rng = np.random.default_rng(9)
    from pyoptimalmotionplanning_b200.backend import CudaSpace
    sp = SpaceSpec([0, 0, -np.inf, -1], [1, 1, np.inf, 1], boxes=[[0.3, 0.3, 0.7, 0.7]], obstacle_dims=(0, 1))
    s = CudaSpace(sp, 0)
    a = rng.random((3000, 4)) * [1, 1, 6, 2] - [0, 0, 0, 1]
    b = rng.random((3000, 4)) * [1, 1, 6, 2] - [0, 0, 0, 1]
    assert (s.edge_check(a, b, 0.01) == oracle.edge_check(sp, a, b, 0.01)).all()
    g = MetricSpec.geodesic_product([("R", 2), ("SO2",), ("R", 1)], [1.0, 0.3, 2.0])
    lens = rng.integers(1, 9, 400)
    off = np.concatenate([[0], np.cumsum(lens)])
    wp = rng.random((off[-1], 4)) * [1, 1, 12, 2] - [0, 0, 3, 1]
    assert (s.edge_check_path(g, wp, off, 0.01) == oracle.edge_check_path(sp, g, wp, off, 0.01)).all()


def test_select_control_batched_matches_oracle():
    from pyoptimalmotionplanning_b200.backend import CudaDynamics
    rng = np.random.default_rng(12)
    cases = [
        (DynamicsSpec(_abi.DYN_DUBINS, 3, 2, _abi.COST_ABS_U0),
         MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi]).with_cost(1.0, 2.0), 64),
        (DynamicsSpec(_abi.DYN_PENDULUM, 2, 2, _abi.COST_TIME, dt=0.01, p=[9.8, 1, 1]),
         MetricSpec.geodesic_product([("SO2",), ("R", 1)], [1.0, 1.0]).with_cost(1.0, None), 64),
        (DynamicsSpec(_abi.DYN_CV, 4, 3, _abi.COST_TIME, dt=0.05), MetricSpec.euclidean(4).with_cost(0.0, None), 33),
    ]
    for dyn, m, S in cases:
        B = 300
        X = dyn.state_dim()
        x = rng.random((B, X))
        xd = rng.random((B, X)) * 1.5
        u = rng.random((B, S, dyn.u_dim)) - 0.5
        u[..., 0] = np.abs(u[..., 0]) * 0.6
        valid = rng.random((B, S)) > 0.1
        valid[5] = False
        best, xb, db = CudaDynamics(dyn).select_control(m, x, xd, u, valid)
        ob, oxb, odb = oracle.select_control(dyn, m, x, xd, u, valid)
        # trig / pow differences can only flip candidates whose distances agree to ~1e-12
        flips = np.nonzero(best != ob)[0]
        for f in flips:
            assert abs(db[f] - odb[f]) <= 1e-11 * max(1.0, abs(odb[f]))
        assert len(flips) <= 2
        same = best == ob
        np.testing.assert_allclose(xb[same], oxb[same], rtol=1e-5, atol=1e-12, equal_nan=True)
        assert best[5] == -1


def test_fused_extend_matches_oracle():
    from pyoptimalmotionplanning_b200.backend import CudaNN, CudaSpace, rrt_extend
    rng = np.random.default_rng(13)
    m = MetricSpec.euclidean(2)
    sp = SpaceSpec([0, 0], [1, 1], boxes=[[0.3, 0.3, 0.7, 0.7], [0.1, 0.6, 0.2, 0.9]])
    s = CudaSpace(sp, 0)
    for n in (1, 37, 5000, 70000):
        pts = rng.random((n, 2))
        nn = CudaNN(m, 0)
        nn.add(pts)
        cur = pts
        for t in range(40):
            xr = rng.random(2)
            parent, xnew, elen, added = rrt_extend(nn, s, xr, 0.15, 0.01, True)
            if not oracle.feasible(sp, [xr])[0]:
                assert parent == -2 and not added
                continue
            op, oxn, oadd = oracle.rrt_extend(m, cur, sp, xr, 0.15, 0.01)
            assert parent == op and added == oadd and (np.array(xnew) == oxn).all()
            assert elen == float(np.sqrt(((cur[op] - oxn) ** 2)[0] + ((cur[op] - oxn) ** 2)[1]))
            if added:
                cur = np.concatenate([cur, oxn[None]])
        assert nn.size() == len(cur)
        idx, _ = nn.nearest(rng.random((50, 2)))   # appended nodes are visible to later queries
        assert idx.max() < len(cur)


def test_errors_are_loud():
    from pyoptimalmotionplanning_b200 import PompError
    from pyoptimalmotionplanning_b200.backend import CudaNN
    nn = CudaNN(MetricSpec.euclidean(2), 0)
    nn.add(np.zeros((4, 2)))
    with pytest.raises(PompError):
        nn.knearest(np.zeros((1, 2)), 129)
    with pytest.raises(PompError):
        nn.remove(10)
    with pytest.raises(PompError):
        nn.set_mask(np.zeros(3, np.uint8))


@pytest.mark.parametrize("k", [1, 4, 16, 32])
@pytest.mark.parametrize("mode", ["strip", "block", "pruned"])
def test_large_batch_2d_paths_match_oracle(k, mode):
    """The large-batch paths (streaming kernel over the strip index -- k = 1 only, larger k falls through to the
    gather kernels -- / fixed-block kernel / pruned traversal) against the oracle on a query subset, and against
    each other on every query."""
    rng = np.random.default_rng(200 + k)
    n, nq = 300000, 40000
    pts = rng.random((n, 2))
    pts[:20000] = 0.5 + 0.002 * rng.standard_normal((20000, 2))     # a dense cluster: tiles overflow the buffer
    q = np.concatenate([rng.random((nq - 3000, 2)), 0.5 + 0.004 * rng.standard_normal((2000, 2)),
                        rng.random((1000, 2)) * 1.2 - 0.1])
    m = MetricSpec.euclidean(2)
    nn = _nn(m, pts, grid=True)
    nn.set_param("strip_search", 1 if mode == "strip" else 0)
    nn.set_param("strip_min_per_tile", 0)
    nn.set_param("block_search", 1 if mode in ("block", "strip") else 0)
    idx, dist, cnt = nn.knearest(q, k)
    sub = rng.choice(nq, 2500, replace=False)
    oi, od, oc = oracle.nn_knearest(m, pts, q[sub], k)
    assert (idx[sub] == oi).all() and (dist[sub] == od).all() and (cnt[sub] == oc).all()
    ref = _nn(m, pts, grid=True)
    ref.set_param("sort_queries", 0)
    ref.set_param("strip_search", 0)
    ref.set_param("block_search", 0)
    ridx, rdist, rcnt = ref.knearest(q, k)
    assert (idx == ridx).all() and (dist == rdist).all()


def test_large_batch_3d_block_path_matches_oracle():
    rng = np.random.default_rng(31)
    n, nq = 250000, 36000
    pts = rng.random((n, 3))
    q = rng.random((nq, 3))
    m = MetricSpec.euclidean(3)
    sub = rng.choice(nq, 1500, replace=False)
    for block in (2, 1):   # 2: fixed-block kernel, 1: pruned traversal (the default in 3-D)
        nn = _nn(m, pts, grid=True)
        nn.set_param("block_search", block)
        for k in (1, 16):
            idx, dist, cnt = nn.knearest(q, k)
            oi, od, oc = oracle.nn_knearest(m, pts, q[sub], k)
            assert (idx[sub] == oi).all() and (dist[sub] == od).all()


def test_tile_path_with_masks_and_tail():
    rng = np.random.default_rng(32)
    m = MetricSpec.euclidean(2)
    pts = rng.random((120000, 2))
    q = rng.random((33000, 2))
    nn = _nn(m, pts[:119000], grid=True)
    nn.build_index()
    nn.add(pts[119000:])
    mask = (rng.random(120000) < 0.2).astype(np.uint8)
    nn.set_mask(mask)
    idx, dist, cnt = nn.knearest(q, 4)
    sub = rng.choice(33000, 1500, replace=False)
    oi, od, oc = oracle.nn_knearest(m, pts, q[sub], 4, mask)
    assert (idx[sub] == oi).all() and (dist[sub] == od).all()


@pytest.mark.parametrize("shape", ["uniform", "clustered", "ragged", "sparse_queries", "ties"])
def test_strip_streaming_kernel_equals_gather_kernels_and_oracle(shape):
    """The persistent bulk-copy kernel over the strip index (strip2d.cuh): uniform data at the bench density,
    clusters that overflow a stage (those tiles go to the overflow list), a grid whose width / height are not
    multiples of the tile, query batches that leave most tiles empty, and a lattice full of exact distance ties
    (lowest insertion index must win, also between the replicated halo copies)."""
    rng = np.random.default_rng(77)
    n, nq = 1 << 20, 1 << 17
    pts = rng.random((n, 2))
    q = rng.random((nq, 2))
    if shape == "clustered":
        pts[:60000] = 0.25 + 0.001 * rng.standard_normal((60000, 2))
        q[:20000] = 0.25 + 0.002 * rng.standard_normal((20000, 2))
    elif shape == "ragged":
        pts = pts * [1.0, 0.37] + [3.0, -2.0]
        q = q * [1.1, 0.45] + [2.95, -2.03]               # some queries outside the bounding box
    elif shape == "sparse_queries":
        q = 0.6 + 0.05 * rng.random((nq, 2))               # a few tiles hold every query
    elif shape == "ties":
        pts = np.round(pts * 1500) / 1500                  # lattice: many duplicates and equidistant pairs
        q = np.round(q * 3000) / 3000
    m = MetricSpec.euclidean(2)
    nn = _nn(m, pts, grid=True)
    nn.set_param("strip_min_per_tile", 0)
    idx, dist = nn.nearest(np.ascontiguousarray(q))
    assert nn.get_stat("strip_launches") > 0
    ref = _nn(m, pts, grid=True)
    ref.set_param("strip_search", 0)
    ridx, rdist = ref.nearest(np.ascontiguousarray(q))
    assert (idx == ridx).all() and (dist == rdist).all()
    sub = rng.choice(nq, 1500, replace=False)
    oi, od = oracle.KDTree(m, pts).nearest(q[sub]) if shape != "ties" else oracle.nn_nearest(m, pts, q[sub[:200]])
    k = len(oi)
    assert (idx[sub[:k]] == oi).all() and (dist[sub[:k]] == od).all()
    # index mutation invalidates the strip copy: appended points must be seen (tail -> gather kernels, then rebuild)
    extra = q[:3000] + 1e-7
    nn.add(extra)
    i2, d2 = nn.nearest(np.ascontiguousarray(q[:40000]))
    assert shape == "ties" or (i2[:3000] >= n).all()     # (on the lattice an old point may coincide with the query)
    oi2, od2 = oracle.KDTree(m, np.concatenate([pts, extra])).nearest(q[:2000]) if shape != "ties" else (i2[:2000], d2[:2000])
    assert (i2[:2000] == oi2).all() and (d2[:2000] == od2).all()


def test_so2_periodic_axis_wraps_and_normalises():
    """SO(2) components are periodic grid axes: neighbours across the 0 / 2 pi seam and angles stored
    outside [0, 2 pi) must come out exactly as the brute-force oracle finds them."""
    m = MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi])
    rng = np.random.default_rng(77)
    n, nq, k = 40000, 600, 4
    pts = rng.random((n, 3)) * [1, 1, 8 * np.pi] - [0, 0, 4 * np.pi]   # angles in [-4 pi, 4 pi)
    q = rng.random((nq, 3)) * [1, 1, 8 * np.pi] - [0, 0, 4 * np.pi]
    q[:200, 2] = rng.choice([0.0, 1e-9, -1e-9, 2 * np.pi, 2 * np.pi - 1e-12, -2 * np.pi], 200)  # on the seam
    for periodic in (1.0, 0.0):
        nn = _nn(m, None, True)
        nn.set_param("periodic_axes", periodic)
        nn.set(pts)
        idx, dist, cnt = nn.knearest(q, k)
        assert nn.get_stat("grid_axes") == (3 if periodic else 2)
        oi, od, oc = oracle.nn_knearest(m, pts, q, k)
        assert (cnt == oc).all()
        assert _near_tie_ok(m, pts, q, idx, oi, od)
        np.testing.assert_allclose(dist, od, rtol=1e-14)
        off, hits = nn.neighbors(q[:150], 0.05)
        ooff, ohits = oracle.nn_neighbors(m, pts, q[:150], 0.05, True)
        assert _hits_equal_mod_boundary(m, pts, q[:150], 0.05, off, hits, ooff, ohits)


def test_rewire_candidates_match_oracle():
    """pomp_rewire_candidates_h (RRT* rewire batch): k nearest of the new node + both straight edges per
    neighbour, against the oracle's knearest and edge checks."""
    from pyoptimalmotionplanning_b200.backend import CudaNN, CudaSpace
    rng = np.random.default_rng(12)
    boxes = np.concatenate([rng.random((60, 2)) * 0.9, np.zeros((60, 2))], 1)
    boxes[:, 2:] = boxes[:, :2] + 0.02 + rng.random((60, 2)) * 0.06
    for D, sp in ((2, SpaceSpec([0, 0], [1, 1], boxes=boxes)),
                  (3, SpaceSpec([0, 0, 0], [1, 1, 1], boxes=np.array([[0.3, 0.3, 0.7, 0.7]])))):
        m = MetricSpec.euclidean(D)
        pts = rng.random((3000, D))
        nn = CudaNN(m, 0)
        nn.set(pts)
        space = CudaSpace(sp, 0)
        for k in (1, 7, 20):
            for _ in range(12):
                x = rng.random(D)
                ids, dist, ok_in, ok_out = nn.rewire_candidates(space, x, k, 0.01)
                oi, od, oc = oracle.nn_knearest(m, pts, [x], k)
                assert len(ids) == int(oc[0]) == k
                assert (ids == oi[0]).all() and (dist == od[0]).all()
                xs = np.repeat(x[None, :], k, 0)
                assert (ok_in == oracle.edge_check(sp, pts[ids], xs, 0.01)).all()
                assert (ok_out == oracle.edge_check(sp, xs, pts[ids], 0.01)).all()
    few = CudaNN(MetricSpec.euclidean(2), 0)
    few.set(pts[:3, :2])
    ids, dist, ok_in, ok_out = few.rewire_candidates(CudaSpace(SpaceSpec([0, 0], [1, 1]), 0), [0.5, 0.5], 8, 0.01)
    assert len(ids) == 3 and ok_in.all() and ok_out.all()


@pytest.mark.parametrize("name", ["euclid2", "se2", "cv"])
@pytest.mark.parametrize("grid", [True, False], ids=["grid", "brute"])
def test_density_matches_oracle(name, grid):
    """pomp_nn_density_h (EST Gaussian density): the summation order differs from the reference's Python loop,
    so the bar is 1e-12 relative (measured ~1e-15) on the sum and exact on the hit count."""
    m = {"euclid2": MetricSpec.euclidean(2), "se2": _metrics()["se2"], "cv": _metrics()["cv"]}[name]
    rng = np.random.default_rng(9)
    hi = np.ones(m.dim)
    if name == "se2":
        hi[2] = 2 * np.pi
    pts = rng.random((20000, m.dim)) * hi
    q = rng.random((300, m.dim)) * hi
    nn = _nn(m, pts, grid)
    r = 0.05 if name != "cv" else 0.2
    dens, cnt = nn.density(q, 4 * r, r, True)
    od, oc = oracle.nn_density(m, pts, q, 4 * r, r, True)
    assert oc.max() > 5
    assert (np.abs(cnt.astype(int) - oc.astype(int)) <= (0 if name == "euclid2" else 1)).all()
    ok = cnt == oc
    np.testing.assert_allclose(dens[ok], od[ok], rtol=1e-12, atol=1e-300)


def test_host_pipeline_graph_replay_sees_new_queries_and_new_points():
    """The pipelined host-buffer path replays CUDA graphs of its chunk searches from the third call on: the
    replays must read the NEW queries, and any change of the node set must drop the captured graphs."""
    rng = np.random.default_rng(5)
    m = MetricSpec.euclidean(2)
    pts = rng.random((300000, 2))
    nn = _nn(m, pts, grid=True)
    nq = 140000      # >= 131072: pipelined over chunks
    sub = rng.choice(nq, 800, replace=False)
    for rep in range(4):
        q = rng.random((nq, 2))
        idx, dist = nn.nearest(q)
        oi, od = oracle.nn_nearest(m, pts, q[sub])
        assert (idx[sub] == oi).all() and (dist[sub] == od).all(), rep
    extra = rng.random((50, 2))
    nn.add(extra)
    pts2 = np.concatenate([pts, extra])
    for rep in range(3):
        q = np.concatenate([extra + 1e-7, rng.random((nq - 50, 2))])
        idx, dist = nn.nearest(q)
        assert (idx[:50] == np.arange(300000, 300050)).all()
        oi, od = oracle.nn_nearest(m, pts2, q[sub])
        assert (idx[sub] == oi).all() and (dist[sub] == od).all(), rep
    nn.remove(300010)
    q = np.concatenate([extra + 1e-7, rng.random((nq - 50, 2))])
    idx, dist = nn.nearest(q)
    assert idx[10] != 300010 and (idx[:10] == np.arange(300000, 300010)).all()


@pytest.mark.parametrize("D", [2, 3, 4, 6])
def test_bruteforce_tiled_nearest_matches_oracle(D):
    """Planner regime (no grid): query batches go through the shared-memory tiled kernel; exact ties (lattice
    points), masks and ragged sizes must give the reference's answer (first minimum in insertion order)."""
    rng = np.random.default_rng(40 + D)
    m = MetricSpec.euclidean(D)
    for n, nq, lattice in ((5000, 1500, False), (7777, 3001, True), (300, 1024, True)):
        pts = rng.integers(0, 6, (n, D)).astype(np.float64) * 0.25 if lattice else rng.random((n, D))
        q = rng.integers(0, 6, (nq, D)).astype(np.float64) * 0.25 + 0.125 if lattice else rng.random((nq, D))
        for tiled in (1, 0):
            nn = _nn(m, pts, grid=False)
            nn.set_param("bf_tiled", tiled)
            idx, dist = nn.nearest(q)
            oi, od = oracle.nn_nearest(m, pts, q)
            assert (idx == oi).all() and (dist == od).all(), (n, nq, lattice, tiled)
            mask = (rng.random(n) < 0.5).astype(np.uint8)
            nn.set_mask(mask)
            idx, dist = nn.nearest(q)
            oi, od = oracle.nn_nearest(m, pts, q, mask)
            assert (idx == oi).all() and (dist == od).all(), (n, nq, lattice, tiled, "masked")


def test_forest_multi_iteration_launches_equal_single_iteration_launches():
    """pomp_forest_extend_multi_h (up to 32 expansions per launch, samples drawn ahead, early stop at the goal with
    the unused random numbers handed back) must grow exactly the trees of the one-expansion-per-launch path."""
    from pyoptimalmotionplanning_b200.rrt import RRTForest, RepeatedRRTForest
    sp = SpaceSpec([0, 0], [1, 1], boxes=[[0.3, 0.3, 0.7, 0.7]])
    seeds = [0, 1, 2, 3, 11]
    a = RRTForest(sp, [0.1, 0.5], seeds, [0.9, 0.5], 0.1, max_step=0.15, resolution=0.01)
    b = RRTForest(sp, [0.1, 0.5], seeds, [0.9, 0.5], 0.1, max_step=0.15, resolution=0.01)
    b.multi = False
    a.planMore(333)          # chunks of 32 + a ragged tail
    b.planMore(333)
    assert a.nodes == b.nodes and a.parents == b.parents and a.costs == b.costs and a.numIters == b.numIters
    ra = RepeatedRRTForest(sp, [0.1, 0.5], seeds, [0.9, 0.5], 0.1, max_step=0.15, resolution=0.01)
    rb = RepeatedRRTForest(sp, [0.1, 0.5], seeds, [0.9, 0.5], 0.1, max_step=0.15, resolution=0.01)
    rb.multi = False
    for _ in range(120):      # the reference harness's loop: trials return at a goal hit, unused samples go back
        ra.planMore(10)
        rb.planMore(10)
        assert ra.bestPathCost == rb.bestPathCost and ra.totalIters == rb.totalIters
    assert ra.nodes == rb.nodes and ra.parents == rb.parents
    assert any(c is not None for c in ra.bestPathCost)


@pytest.mark.parametrize("k", [1, 5, 16, 32])
def test_se2_compile_time_kernel_equals_generic_kernel_and_oracle(k):
    """grid_knn_se2_kernel performs metric_eval's operations in the same order: its answers must be IDENTICAL to
    the generic kernel's (bit-equal distances, same indices), masks / tails / seam angles included; against the
    oracle the usual libm-pow tolerance of the product metrics applies."""
    m = MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi])
    rng = np.random.default_rng(100 + k)
    n, nq = 60000, 3000
    pts = rng.random((n, 3)) * [1, 1, 8 * np.pi] - [0, 0, 4 * np.pi]
    q = rng.random((nq, 3)) * [1, 1, 8 * np.pi] - [0, 0, 4 * np.pi]
    q[:300, 2] = rng.choice([0.0, 1e-9, -1e-9, 2 * np.pi, 2 * np.pi - 1e-12, -2 * np.pi], 300)
    q[300:600] = pts[rng.integers(0, n, 300)]
    mask = (rng.random(n) < 0.3).astype(np.uint8)
    res = {}
    for fast in (1, 0):
        nn = _nn(m, None, True)
        nn.set_param("se2_fast", fast)
        nn.set(pts[:n - 700])
        nn.build_index()
        nn.add(pts[n - 700:])          # unindexed tail
        out = [nn.knearest(q, k)]
        nn.set_mask(mask)
        out.append(nn.knearest(q, k))
        res[fast] = out
    for a, b in zip(res[1], res[0]):
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all()
    sub = rng.choice(nq, 400, replace=False)
    oi, od, oc = oracle.nn_knearest(m, pts, q[sub], k)
    idx, dist, cnt = res[1][0]
    assert (cnt[sub] == oc).all()
    assert _near_tie_ok(m, pts, q[sub], idx[sub], oi, od)
    np.testing.assert_allclose(dist[sub], od, rtol=1e-14)


def _numpy_merge(idx_in, dist_in, k):
    """(distance, global index) merge of G candidate lists per query -- the rule of SURVEY 8e."""
    G, nq, _ = idx_in.shape
    ii = np.transpose(idx_in, (1, 0, 2)).reshape(nq, -1)
    dd = np.transpose(dist_in, (1, 0, 2)).reshape(nq, -1).copy()
    dd[ii < 0] = np.inf
    key_i = np.where(ii < 0, np.iinfo(np.int64).max, ii)
    order = np.lexsort((key_i, dd), axis=1)[:, :k]
    oi = np.take_along_axis(ii, order, 1)
    od = np.take_along_axis(dd, order, 1)
    oi[np.isinf(od)] = -1
    return oi, od


@pytest.mark.parametrize("k", [1, 16, 40, 100])
@pytest.mark.parametrize("G", [2, 3, 8])
def test_merge_candidates_kernel_matches_lexsort_merge(G, k):
    """pomp_nn_merge_candidates (the kernel behind the NCCL node-sharded path) against the NumPy merge the gloo
    test uses, including short lists (-1 / +inf padding) and exact cross-shard distance ties."""
    import torch
    from pyoptimalmotionplanning_b200.backend import merge_candidates_device
    rng = np.random.default_rng(1000 * G + k)
    nq = 3000
    dist_in = np.sort(rng.random((G, nq, k)), axis=2)
    dist_in[:, :500] = np.round(dist_in[:, :500] * 8) / 8          # many exact ties across shards
    dist_in = np.sort(dist_in, axis=2)
    idx_in = rng.permutation(G * nq * k).reshape(G, nq, k).astype(np.int64)
    # within a shard, equal distances must come in index order (that is what a shard's k-NN returns)
    for g in range(G):
        order = np.lexsort((idx_in[g], dist_in[g]), axis=1)
        idx_in[g] = np.take_along_axis(idx_in[g], order, 1)
        dist_in[g] = np.take_along_axis(dist_in[g], order, 1)
    short = rng.integers(0, k + 1, (G, nq))
    short[:, 1000:] = k
    for g in range(G):
        pad = np.arange(k)[None, :] >= short[g][:, None]
        idx_in[g][pad] = -1
        dist_in[g][pad] = np.inf
    ti, td = torch.from_numpy(idx_in).cuda(), torch.from_numpy(dist_in).cuda()
    oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    od = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    oc = torch.empty(nq, dtype=torch.int32, device="cuda")
    merge_candidates_device(ti.data_ptr(), td.data_ptr(), G, nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    wi, wd = _numpy_merge(idx_in, dist_in, k)
    assert (oi.cpu().numpy() == wi).all()
    assert (od.cpu().numpy() == wd).all()
    assert (oc.cpu().numpy() == (wi >= 0).sum(1)).all()


def test_node_sharded_answer_equals_single_index_on_one_gpu():
    """The node-sharded k-NN pipeline of sharded.py with its collectives replaced by slicing (one process, G
    shards on one GPU): local top-k per shard + pomp_nn_merge_candidates == one index over all the points."""
    import torch
    from pyoptimalmotionplanning_b200.backend import CudaNN, merge_candidates_device
    rng = np.random.default_rng(5)
    G, n, nq, k = 4, 120000, 40000, 8
    pts = rng.random((G * n, 2))
    pts[n + 5] = pts[7]                                   # cross-shard duplicate: tie broken by global index
    q = rng.random((nq, 2))
    m = MetricSpec.euclidean(2)
    whole = _nn(m, pts, grid=True)
    wi, wd, _ = whole.knearest(q, k)
    st = torch.cuda.current_stream().cuda_stream
    qd = torch.from_numpy(q).cuda()
    ci = torch.empty((G, nq, k), dtype=torch.int64, device="cuda")
    cd = torch.empty((G, nq, k), dtype=torch.float64, device="cuda")
    shards = []
    for g in range(G):
        nn = CudaNN(m, 0)
        nn.set(pts[g * n:(g + 1) * n])
        nn.knearest_device(qd.data_ptr(), nq, k, ci[g].data_ptr(), cd[g].data_ptr(), None, st)
        shards.append(nn)
        ci[g] = torch.where(ci[g] >= 0, ci[g] + g * n, ci[g])
    oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    od = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    merge_candidates_device(ci.data_ptr(), cd.data_ptr(), G, nq, k, oi.data_ptr(), od.data_ptr(), None, st)
    torch.cuda.synchronize()
    assert (oi.cpu().numpy() == wi).all() and (od.cpu().numpy() == wd).all()


def test_graph_replay_survives_interleaved_batch_sizes():
    """the pipelined host call replays captured CUDA graphs that bake scratch pointers in.  A batch of
    another size (or a radius call) in between may regrow that scratch: the graphs must be recaptured, not
    replayed on freed memory."""
    rng = np.random.default_rng(41)
    m = MetricSpec.euclidean(2)
    pts = rng.random((400000, 2))
    nn = _nn(m, pts, grid=True)
    big = rng.random((200000, 2))
    mid = rng.random((120000, 2))
    kd = oracle.KDTree(m, pts)
    sub = rng.choice(200000, 2000, replace=False)
    want_i, want_d = kd.nearest(big[sub])
    for _ in range(3):                       # direct run, capture, replay
        i, d = nn.nearest(big)
        assert (i[sub] == want_i).all() and (d[sub] == want_d).all()
    mi, md = nn.nearest(mid)                 # single-pass size between the chunk sizes: regrows scratch
    oi, od = kd.nearest(mid[:2000])
    assert (mi[:2000] == oi).all() and (md[:2000] == od).all()
    nn.neighbors(mid[:50000], 0.002)
    nn.knearest(rng.random((300000, 2)), 4)
    for _ in range(3):
        i, d = nn.nearest(big)
        assert (i[sub] == want_i).all() and (d[sub] == want_d).all()
        full_i, full_d = i, d
    ref = _nn(m, pts, grid=True)
    ref.set_param("use_graphs", 0)
    ri, rd = ref.nearest(big)
    assert (full_i == ri).all() and (full_d == rd).all()


def test_spatial_shard_kernels_match_numpy_ops():
    """csrc/shard.cu (route / finalize / unpack) against the NumPy stand-ins the gloo tests use, then the whole
    SpatialShardedNearestNeighbors on CUDA tensors in a one-rank NCCL group against the oracle."""
    import socket
    import torch
    import torch.distributed as dist
    from pyoptimalmotionplanning_b200.sharded import _CudaShardOps, SpatialShardedNearestNeighbors
    from tests.test_sharded_gloo import NumpyShardOps
    rng = np.random.default_rng(8)
    m = MetricSpec.euclidean(2)
    cu, nu = _CudaShardOps(m, 0), NumpyShardOps(m)
    q = rng.random((50000, 2))
    q[:5, 1] = [0.25, 0.5, 0.75, np.nan, -3.0]            # exactly on slab faces, NaN, far outside
    bounds = [0.25, 0.5, 0.75]
    qs, perm, counts = cu.route(torch.from_numpy(q).cuda(), 1, bounds, 4)
    nqs, nperm, ncounts = nu.route(torch.from_numpy(q), 1, bounds, 4)
    assert counts[:3] == ncounts[:3] or np.isnan(q[:, 1]).any()
    perm_h, qs_h = perm.cpu().numpy(), qs.cpu().numpy()
    assert sorted(perm_h.tolist()) == list(range(len(q)))
    assert ((qs_h == q[perm_h]) | np.isnan(qs_h)).all()
    start = 0
    for r, c in enumerate(counts):                        # every routed block holds exactly its slab's queries
        y = qs_h[start:start + c, 1]
        lo = -np.inf if r == 0 else bounds[r - 1]
        hi = np.inf if r == 3 else bounds[r]
        assert (np.isnan(y) | ((y >= lo) & (y < hi))).all()
        start += c
    pts = rng.random((30000, 2))
    cu.set_points(torch.from_numpy(pts).cuda())
    nu.set_points(torch.from_numpy(pts))
    l2g = torch.from_numpy(np.sort(rng.choice(10 ** 6, 30000, replace=False)))
    halo = torch.from_numpy((rng.random(30000) < 0.2).astype(np.uint8))
    qq = torch.from_numpy(rng.random((3000, 2)))
    for drop in (False, True):
        il, dl = cu.local_knn(qq.cuda(), 5)
        rec, fl = cu.finalize(qq.cuda(), 1, 0.1, 0.9, 5, il, dl, l2g.cuda(), halo.cuda(), drop, True)
        nil, ndl = nu.local_knn(qq, 5)
        nrec, nfl = nu.finalize(qq, 1, 0.1, 0.9, 5, nil, ndl, l2g, halo, drop, True)
        assert (rec.cpu() == nrec).all() and (fl.cpu() == nfl).all()
        p = torch.from_numpy(rng.permutation(3000))
        ui, ud = cu.unpack(rec, p.cuda(), 3000, 5)
        ni, nd = nu.unpack(nrec, p, 3000, 5)
        assert (ui.cpu() == ni).all() and (ud.cpu() == nd).all()
    # the whole object, one rank over NCCL
    if not dist.is_initialized():
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    big = rng.random((200000, 2))
    sp = SpatialShardedNearestNeighbors(m, device=0)
    sp.set(torch.from_numpy(big).cuda(), 0)
    bq = rng.random((40000, 2)) * 1.1 - 0.05
    idx, dst, st = sp.knearest(torch.from_numpy(bq).cuda(), 4)
    oi, od, _ = oracle.KDTree(m, big).knearest(bq, 4)
    assert (idx.cpu().numpy() == oi).all() and (dst.cpu().numpy() == od).all()
    fi, fd, fst = sp.knearest_fast(torch.from_numpy(bq).cuda(), 4)            # fixed-capacity, sync-free variant
    assert (fi.cpu().numpy() == oi).all() and (fd.cpu().numpy() == od).all() and "deferred" in fst
    sp.hi_cov = 0.97      # pretend the slab ends there: queries beyond lose their certificate -> exact fallback
    fi, fd, fst = sp.knearest_fast(torch.from_numpy(bq).cuda(), 4)
    assert (fi.cpu().numpy() == oi).all() and (fd.cpu().numpy() == od).all() and fst["deferred"] > 500
    fi, fd, fst = sp.knearest_fast(torch.from_numpy(bq).cuda(), 4, slack=0.5)   # blocks too small: the rest is deferred
    assert (fi.cpu().numpy() == oi).all() and (fd.cpu().numpy() == od).all() and fst["deferred"] > 10000
    sp.hi_cov = float("inf")
    # peer-memory exchange (NVLink stores from the routing / finalize kernels; one rank: its own buffers), synchronous
    # and streamed (wait() of batch i after batch i + 1 was issued), with and without deferred queries
    try:
        sp.knearest_p2p(torch.from_numpy(bq).cuda(), 4)
        p2p = True
    except Exception as e:      # no symmetric memory on this box
        print("peer-memory exchange unavailable:", str(e)[:200])
        p2p = False
    if p2p:
        qd = torch.from_numpy(bq).cuda()
        qd2 = torch.from_numpy(bq[::-1].copy()).cuda()
        pi_, pd_, pst = sp.knearest_p2p(qd, 4)
        assert (pi_.cpu().numpy() == oi).all() and (pd_.cpu().numpy() == od).all() and pst["deferred"] == 0
        pend = [sp.knearest_p2p_async(qd, 4), sp.knearest_p2p_async(qd2, 4)]
        sp.hi_cov = 0.97                                  # the third batch loses certificates -> fallback in wait()
        pend.append(sp.knearest_p2p_async(qd, 4))
        sp.hi_cov = float("inf")
        pend.append(sp.knearest_p2p_async(qd, 4, slack=0.5))   # blocks too small -> deferred at routing time
        res = [p_.wait() for p_ in pend]
        for j, (ri, rd, rst) in enumerate(res):
            want_i, want_d = (oi[::-1], od[::-1]) if j == 1 else (oi, od)
            assert (ri.cpu().numpy() == want_i).all() and (rd.cpu().numpy() == want_d).all(), j
        assert res[0][2]["deferred"] == 0 and res[2][2]["deferred"] > 500 and res[3][2]["deferred"] > 10000
    off, hits = sp.neighbors(torch.from_numpy(bq[:2000]).cuda(), 0.004)
    ooff, ohits = oracle.nn_neighbors(m, big, bq[:2000], 0.004, True)
    assert (off.cpu().numpy() == ooff).all() and (hits.cpu().numpy() == ohits).all()
    dist.destroy_process_group()


# ------------------------------------------------------------------ warp-per-query shell kernel (grid_warp.cuh)
def _lexsort_knn(pts, q, k, mask=None):
    """NumPy brute force with the library's tie rule: the k smallest (distance, insertion index) keys; distances
    summed in coordinate order like vectorops.distance (vectorops.py:94-102)."""
    s = np.zeros((len(q), len(pts)))
    for j in range(pts.shape[1]):
        t = pts[None, :, j] - q[:, None, j]
        s = s + t * t
    dd = np.sqrt(s)
    if mask is not None:
        dd = np.where(mask[None, :] != 0, np.inf, dd)
    order = np.lexsort((np.broadcast_to(np.arange(len(pts)), dd.shape), dd), axis=1)[:, :k]
    od = np.take_along_axis(dd, order, axis=1)
    oi = np.where(np.isfinite(od), order, -1)
    return oi, od


@pytest.mark.parametrize("lanes", [0, 8, 16, 32])
@pytest.mark.parametrize("D", [2, 3, 4, 6])
@pytest.mark.parametrize("k", [1, 5, 16, 32])
def test_warp_shell_kernel_matches_oracle(D, k, lanes):
    """warp_search = 2 forces the warp-per-query kernel for every Euclidean (D, k): uniform cloud, queries outside
    the bounding box, sorted (large) and unsorted (small) batches, several first-shell radii (too small -> radius
    doubling and second shells; too large -> one fat shell)."""
    rng = np.random.default_rng(500 + 10 * D + k)
    n = 150000
    pts = rng.random((n, D))
    q = np.concatenate([rng.random((1900, D)), rng.random((100, D)) * 1.6 - 0.3])
    m = MetricSpec.euclidean(D)
    oi, od, oc = oracle.nn_knearest(m, pts, q, k)
    for need in (0, 0.05, 40 * k):
        nn = _nn(m, pts, grid=True)
        nn.set_param("warp_search", 2)
        nn.set_param("warp_lanes", lanes)
        nn.set_param("warp_need", need)
        nn.set_param("sort_min_queries", 1000)
        idx, dist, cnt = nn.knearest(q, k)
        assert (cnt == oc).all(), need
        assert (idx == oi).all(), need
        assert (dist == od).all(), need
        idx, dist, cnt = nn.knearest(q[:300], k)          # below the sort threshold: qperm = NULL
        assert (idx == oi[:300]).all() and (dist == od[:300]).all()
    if k == 1:
        i1, d1 = nn.nearest(q)
        assert (i1 == oi[:, 0]).all() and (d1 == od[:, 0]).all()


@pytest.mark.parametrize("lanes", [0, 8, 32])
@pytest.mark.parametrize("D", [2, 3, 6])
def test_warp_shell_kernel_masks_tail_ties_and_clusters(D, lanes):
    rng = np.random.default_rng(77 + D)
    m = MetricSpec.euclidean(D)
    # (1) tombstones + reject mask + unindexed tail; more than half of the points masked in one corner, so lists
    # stay short after the first shells
    pts = rng.random((40000, D))
    q = rng.random((600, D))
    nn = _nn(m, pts[:38000], grid=True)
    nn.set_param("warp_search", 2)
    nn.set_param("warp_lanes", lanes)
    nn.build_index()
    nn.add(pts[38000:])
    mask = np.zeros(40000, np.uint8)
    mask[(pts[:, 0] < 0.6) & (pts[:, 1] < 0.7)] = 1
    mask[rng.choice(40000, 3000, replace=False)] = 1
    nn.set_mask(mask)
    for k in (1, 7, 32):
        idx, dist, cnt = nn.knearest(q, k)
        oi, od, oc = oracle.nn_knearest(m, pts, q, k, mask)
        assert (idx == oi).all() and (dist == od).all() and (cnt == oc).all(), k
    # (2) lattice with duplicates: exact ties everywhere, the lowest insertion indices win
    side = {2: 40, 3: 12, 6: 3}[D]
    g = np.stack(np.meshgrid(*([np.arange(side) / side] * D), indexing="ij"), -1).reshape(-1, D)
    lat = np.concatenate([g, g[::3], g[::5]])
    ql = np.concatenate([g[::7] + 0.5 / side * 0.5, g[::11], rng.random((50, D))])
    nn = _nn(m, lat, grid=True)
    nn.set_param("warp_search", 2)
    nn.set_param("warp_lanes", lanes)
    for k in (1, 6, 20):
        oi, od = _lexsort_knn(lat, ql, k)
        idx, dist, cnt = nn.knearest(ql, k)
        assert (dist == od).all(), k
        assert (idx == oi).all(), k
    # (3) clustered cloud (the global density is useless there), fewer points than k, non-finite queries
    centres = rng.random((20, D))
    clu = centres[rng.integers(0, 20, 60000)] + rng.normal(0, 0.002, (60000, D))
    qc = np.concatenate([clu[::200] + rng.normal(0, 0.001, (300, D)), rng.random((100, D))])
    nn = _nn(m, clu, grid=True)
    nn.set_param("warp_search", 2)
    nn.set_param("warp_lanes", lanes)
    for k in (1, 16):
        oi, od, oc = oracle.nn_knearest(m, clu, qc, k)
        idx, dist, cnt = nn.knearest(qc, k)
        assert (idx == oi).all() and (dist == od).all(), k
    few = rng.random((20, D))
    nn = _nn(m, few, grid=True)
    nn.set_param("warp_search", 2)
    nn.set_param("warp_lanes", lanes)
    qf = rng.random((70, D))
    qf[3, 0] = np.nan
    qf[5, D - 1] = np.inf
    idx, dist, cnt = nn.knearest(qf, 32)
    oi, od, oc = oracle.nn_knearest(m, few, qf, 32)
    assert (idx == oi).all() and (cnt == oc).all()
    assert cnt[3] == 0 and cnt[5] == 0 and (cnt[[0, 1, 2]] == 20).all()


def test_single_radius_query_fast_path_matches_oracle():
    """One neighbors() query in the brute-force regime: the one-launch path (hits appended in mapped host memory,
    sorted on the host) against the oracle and against the general count / scan / fill path -- inclusive and
    exclusive radius, generic metric, masks, and a query with more hits than the box holds (falls back)."""
    rng = np.random.default_rng(31)
    for m, D in ((MetricSpec.euclidean(2), 2), (_metrics()["se2"], 3), (MetricSpec.l1(3), 3)):
        pts = rng.random((20000, D))
        nn = _nn(m, pts)
        ref = _nn(m, pts)
        ref.set_param("single_radius", 0)
        mask = (rng.random(20000) < 0.3).astype(np.uint8)
        for masked in (False, True):
            nn.set_mask(mask if masked else None)
            ref.set_mask(mask if masked else None)
            for r in (0.0, 0.02, 0.3, 5.0):                  # 5.0: every unmasked point (> 8192 hits)
                for incl in (True, False):
                    q = rng.random((1, D))
                    off, hits = nn.neighbors(q, r, incl)
                    roff, rhits = ref.neighbors(q, r, incl)
                    assert (off == roff).all() and (hits == rhits).all(), (D, masked, r, incl)
                    if m.kind != _abi.METRIC_GEODESIC_PRODUCT:     # product metrics: 2-ulp boundary, see above
                        ooff, ohits = oracle.nn_neighbors(m, pts, q, r, incl, mask if masked else None)
                        assert (off == ooff).all() and (hits == ohits).all(), (D, masked, r, incl)
    # a point exactly at the radius: kd-tree semantics (d <= r) vs brute-force semantics (d < r)
    nn = _nn(MetricSpec.euclidean(2), np.array([[0.0, 0.0], [3.0, 4.0], [6.0, 8.0]]))
    assert nn.neighbors(np.array([[0.0, 0.0]]), 5.0, True)[1].tolist() == [0, 1]
    assert nn.neighbors(np.array([[0.0, 0.0]]), 5.0, False)[1].tolist() == [0]
