"""CUDA path vs the reference-generated golden fixtures (tests/golden), through the C-ABI."""
import math
import random

import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _ulps(a, b):
    if a == b:
        return 0
    if math.isinf(a) or math.isinf(b) or math.isnan(a) or math.isnan(b):
        return 1 << 62
    ia = np.float64(a).view(np.int64)
    ib = np.float64(b).view(np.int64)
    return abs(int(ia) - int(ib))


# metrics whose reference form goes through libm pow(x, 2.0) (not correctly rounded in glibc);
# the device squares with x*x, so distances may differ in the last place.
POW_METRICS = ("weighted", "product")


@pytest.mark.parametrize("case", G.load("metrics.json"), ids=lambda c: c["name"])
def test_metric_values(case):
    from pyoptimalmotionplanning_b200.backend import CudaNN
    m = G.metric_spec(case["spec"])
    nn = CudaNN(m, 0)
    inexact = 0
    for a, b, d in zip(case["a"], case["b"], case["d"]):
        nn.reset()
        nn.add_one(a)
        idx, got = nn.nearest_one(b)
        if math.isinf(d):
            assert idx == -1 and math.isinf(got)
            continue
        assert idx == 0
        u = _ulps(got, d)
        if case["spec"]["kind"] in POW_METRICS:
            assert u <= 2, (case["name"], a, b, got, d)
            inexact += (u != 0)
        else:
            assert u == 0, (case["name"], a, b, got, d)
    # pow-vs-multiply disagreements are rare (measured 0.086 % per squaring)
    assert inexact <= max(3, len(case["d"]) // 20)


@pytest.mark.parametrize("force_grid", [False, True], ids=["brute", "grid"])
@pytest.mark.parametrize("case", G.load("nn.json"), ids=lambda c: c["name"])
def test_nn_golden(case, force_grid):
    from pyoptimalmotionplanning_b200.backend import CudaNN
    m = G.metric_spec(case["spec"])
    pts, q = np.array(case["pts"]), np.array(case["q"])
    nn = CudaNN(m, 0)
    if force_grid:
        nn.set_param("grid_min_points", 1)
        nn.set_param("grid_min_queries", 1)
    nn.set(pts)
    idx, _ = nn.nearest(q)
    assert idx.tolist() == case["nearest_brute"]
    kidx, _, cnt = nn.knearest(q, 7)
    assert [r[:c].tolist() for r, c in zip(kidx, cnt)] == case["knearest_kd"]
    off, hits = nn.neighbors(q, case["radius"], inclusive=False)
    assert [hits[off[i]:off[i + 1]].tolist() for i in range(len(q))] == case["neighbors_brute_lt"]
    off, hits = nn.neighbors(q, case["radius"], inclusive=True)
    assert [hits[off[i]:off[i + 1]].tolist() for i in range(len(q))] == case["neighbors_kd_le"]
    mask = np.zeros(len(pts), np.uint8)
    mask[case["rejected"]] = 1
    nn.set_mask(mask)
    idx, _ = nn.nearest(q)
    assert idx.tolist() == case["nearest_brute_filtered"]
    kidx, _, cnt = nn.knearest(q, 7)
    assert [r[:c].tolist() for r, c in zip(kidx, cnt)] == case["knearest_kd_filtered"]
    nn.set_mask(None)
    # single-query fast path
    for j in range(0, len(q), 7):
        i1, _ = nn.nearest_one(q[j])
        assert i1 == case["nearest_brute"][j]


@pytest.mark.parametrize("case", G.load("edges.json"), ids=lambda c: c["name"])
def test_feasible_and_edges_golden(case):
    from pyoptimalmotionplanning_b200.backend import CudaSpace
    s = CudaSpace(G.space_spec(case["space"]), 0)
    assert s.feasible(case["pts"]).tolist() == case["feasible"]
    for res, ok in case["edge_ok"].items():
        assert s.edge_check(case["a"], case["b"], float(res)).tolist() == ok, res
    for j in range(0, len(case["a"]), 11):
        assert s.edge_check_one(case["a"][j], case["b"][j], 0.01) == case["edge_ok"]["0.01"][j]


@pytest.mark.parametrize("case", G.load("dynamics.json"), ids=lambda c: c["name"])
def test_next_state_golden(case):
    from pyoptimalmotionplanning_b200.backend import CudaDynamics
    dyn = G.dyn_spec(case["spec"])
    got = CudaDynamics(dyn).next_state(case["x"], case["u"])
    want = np.array(case["xnext"])
    # north star: states within 1e-5 relative tolerance (sin/cos and x**2 differ from libm in the last ulp)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-12)
    if case["name"] in ("adaptor_cost",):
        assert (got == want).all()


@pytest.mark.parametrize("case", G.load("control_edges.json"), ids=lambda c: c["name"])
def test_control_edges_golden(case):
    from pyoptimalmotionplanning_b200.backend import CudaSpace
    s = CudaSpace(G.space_spec(case["space"]), 0)
    dyn = G.dyn_spec(case["dyn"])
    g = G.metric_spec(case["geodesic"])
    got = s.edge_check_control(dyn, g, case["x"], case["u"], case["res"])
    assert got.tolist() == case["ok"]
    if "paths" in case:
        paths = case["paths"]
        off = np.cumsum([0] + [len(p) for p in paths])
        wp = np.concatenate([np.array(p).reshape(-1, s.dim) for p in paths])
        got = s.edge_check_path(g, wp, off, case["res"])
        assert got.tolist() == case["ok"][:len(paths)]


@pytest.mark.parametrize("case", G.load("select_control.json"), ids=lambda c: c["name"])
def test_select_control_golden(case):
    from pyoptimalmotionplanning_b200.backend import CudaDynamics
    dyn = G.dyn_spec(case["dyn"])
    m = G.metric_spec(case["metric"])
    rows = case["rows"]
    best, xb, db = CudaDynamics(dyn).select_control(m, [r["x"] for r in rows], [r["xdes"] for r in rows],
                                                    [r["u"] for r in rows], [r["valid"] for r in rows])
    assert best.tolist() == [r["best"] for r in rows]


@pytest.mark.parametrize("name", ["bugtrap", "kink"])
def test_rrt_tree_identical_to_reference(name):
    """G2: plain RRT exploration tree, 2000 iterations, seed 12345 -- identical nodes, parents and
    roadmap hash to the reference planner."""
    from pyoptimalmotionplanning_b200.rrt import RRT
    g = G.load("planners.json")["G2_%s_bruteforce" % name]
    random.seed(g["seed"])
    rrt = RRT(G.space_spec(g["space"]), g["start"], None, None, max_step=0.15, resolution=0.01)
    rrt.planMore(g["iters"])
    assert len(rrt.nodes) == g["n"]
    assert rrt.parents == g["parents"]
    assert rrt.nodes == g["nodes"]
    assert rrt.roadmap_hash() == g["hash"]
    assert rrt.roadmap_hash() == G.load("planners.json")["G2_%s_kdtree" % name]["hash"]


def test_rrrt_bugtrap_cost_trace_identical_to_reference():
    """G1: r-rrt on Bugtrap, seed 0: the (iterations, bestPathCost) change points are identical."""
    from pyoptimalmotionplanning_b200.rrt import RepeatedRRT
    g = G.load("planners.json")["G1_rrrt_bugtrap_bruteforce"]
    random.seed(g["seed"])
    p = RepeatedRRT(G.space_spec(g["space"]), g["start"], g["goal"], g["goal_radius"], max_step=0.15,
                    resolution=0.01)
    cps, last, it = [], None, 0
    while it < g["iters"]:
        p.planMore(10)
        it += 10
        if p.bestPathCost != last:
            cps.append([it, p.bestPathCost])
            last = p.bestPathCost
    assert cps == g["change_points"]


def test_forest_trials_identical_to_reference_trees():
    """Lock-step multi-trial kernel: every trial grows the reference planner's tree of its own seed
    (G2 for the golden seed, and the single-trial CUDA driver for the others)."""
    from pyoptimalmotionplanning_b200.rrt import RRT, RRTForest
    for name in ("bugtrap", "kink"):
        g = G.load("planners.json")["G2_%s_bruteforce" % name]
        seeds = [3, g["seed"], 11, 12, 13]
        f = RRTForest(G.space_spec(g["space"]), g["start"], seeds, None, None, max_step=0.15, resolution=0.01)
        f.planMore(g["iters"])
        v = f.trial(1)
        assert v.nodes == g["nodes"] and v.parents == g["parents"] and v.roadmap_hash() == g["hash"]
        random.seed(seeds[3])
        single = RRT(G.space_spec(g["space"]), g["start"], None, None, max_step=0.15, resolution=0.01)
        single.planMore(g["iters"])
        assert f.trial(3).nodes == single.nodes and f.trial(3).parents == single.parents


def test_rrrt_forest_cost_traces_identical_to_reference():
    """r-rrt on Bugtrap, 10 seeded trials in lock step (main.py's numTrials): trial 0 reproduces G1."""
    from pyoptimalmotionplanning_b200.rrt import RepeatedRRTForest
    g = G.load("planners.json")["G1_rrrt_bugtrap_bruteforce"]
    rf = RepeatedRRTForest(G.space_spec(g["space"]), g["start"], list(range(10)), g["goal"], g["goal_radius"],
                           max_step=0.15, resolution=0.01)
    seen, last = [], None
    for it in range(g["iters"]):
        rf.planMore(1)
        if rf.bestPathCost[0] != last:
            last = rf.bestPathCost[0]
            seen.append(last)
    want = [c[1] for c in g["change_points"]]
    assert seen[-1] == want[-1]
    assert [c for c in want if c in seen] == want        # every sampled golden cost was passed through
    assert all(c is not None for c in rf.bestPathCost)   # all ten trials found paths
