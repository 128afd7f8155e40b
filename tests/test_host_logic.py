"""Host-side logic of the drop-ins, CPU only.  The CUDA back end is replaced by the oracle-backed
fake (tests/fake_backend.py) so that index bookkeeping, filters, metric refresh, interpolator
dispatch and RNG order can be checked here -- against the UNMODIFIED reference where its checkout
is present (it is not on the GPU box; those tests then skip)."""
import hashlib
import random

import numpy as np
import pytest

import pyoptimalmotionplanning_b200 as P
from pyoptimalmotionplanning_b200 import _abi
from pyoptimalmotionplanning_b200.descriptors import metric_from_pomp, space_from_pomp, dynamics_from_pomp
from oracle import refshim
from tests import golden_util as G
from tests.fake_backend import OracleNN, OracleSpace, OracleDynamics, OracleForest, oracle_rrt_extend

needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference checkout not present")


def fake_nn(metric, method="b200"):
    return P.NearestNeighbors(metric, method, _backend_factory=lambda spec, dev: OracleNN(spec, dev))


def fake_checker(space, res):
    return P.EpsilonEdgeChecker(space, res, _backend_factory=lambda spec, dev: OracleSpace(spec, dev))


def roadmap_hash(V, E):
    return hashlib.sha256(repr((V, [(i, j) for i, j, u in E])).encode()).hexdigest()[:16]


# ------------------------------------------------------------------ driver on goldens (no reference needed)
@pytest.mark.parametrize("name", ["bugtrap", "kink"])
def test_rrt_driver_reproduces_reference_tree(name):
    from pyoptimalmotionplanning_b200.rrt import RRT
    g = G.load("planners.json")["G2_%s_bruteforce" % name]
    random.seed(g["seed"])
    rrt = RRT(G.space_spec(g["space"]), g["start"], None, None, max_step=0.15, resolution=0.01,
              _backends=(OracleNN, OracleSpace, oracle_rrt_extend))
    rrt.planMore(g["iters"])
    assert rrt.parents == g["parents"] and rrt.nodes == g["nodes"] and rrt.roadmap_hash() == g["hash"]


def test_rrrt_driver_reproduces_reference_costs():
    from pyoptimalmotionplanning_b200.rrt import RepeatedRRT
    g = G.load("planners.json")["G1_rrrt_bugtrap_bruteforce"]
    random.seed(g["seed"])
    p = RepeatedRRT(G.space_spec(g["space"]), g["start"], g["goal"], g["goal_radius"], max_step=0.15,
                    resolution=0.01, _backends=(OracleNN, OracleSpace, oracle_rrt_extend))
    cps, last, it = [], None, 0
    while it < 1500:
        p.planMore(10)
        it += 10
        if p.bestPathCost != last:
            cps.append([it, p.bestPathCost])
            last = p.bestPathCost
    want = [c for c in g["change_points"] if c[0] <= 1500]
    assert cps == want


def test_forest_driver_trials_equal_single_trial_goldens():
    """Lock-step multi-trial driver: trial t under seeds[t] grows the single-trial tree of that seed."""
    from pyoptimalmotionplanning_b200.rrt import RRTForest, RepeatedRRTForest
    g = G.load("planners.json")["G2_bugtrap_bruteforce"]
    f = RRTForest(G.space_spec(g["space"]), g["start"], [7, g["seed"], 99], None, None, max_step=0.15,
                  resolution=0.01, _backends=(OracleForest, OracleSpace))
    f.planMore(600)
    v = f.trial(1)
    n = len(v.nodes)
    assert n > 300 and v.nodes == g["nodes"][:n] and v.parents == g["parents"][:n]
    assert f.trial(0).nodes != v.nodes
    g1 = G.load("planners.json")["G1_rrrt_bugtrap_bruteforce"]
    rf = RepeatedRRTForest(G.space_spec(g1["space"]), g1["start"], [0, 1], g1["goal"], g1["goal_radius"],
                           max_step=0.15, resolution=0.01, _backends=(OracleForest, OracleSpace))
    cps, last = [], None
    for it in range(1, 701):
        rf.planMore(1)
        if rf.bestPathCost[0] != last:
            last = rf.bestPathCost[0]
            cps.append([-(-it // 10) * 10, last])   # the golden samples the cost every 10 iterations
    want = [c for c in g1["change_points"] if c[0] <= 700]
    assert [c[1] for c in cps if c[1] is not None][-1] == want[-1][1]
    assert sorted(set(c[1] for c in cps if c[1] is not None), reverse=True) == [c[1] for c in want]


def test_forest_results_csv_has_the_reference_layout(tmp_path):
    """testForest writes pomp/planners/test.py's CSV (header, a row per cost change, a closing row per trial);
    trial 0 under seed 0 reproduces the golden cost trace of the reference's r-rrt run."""
    import csv
    from pyoptimalmotionplanning_b200.rrt import RepeatedRRTForest, testForest
    g1 = G.load("planners.json")["G1_rrrt_bugtrap_bruteforce"]
    rf = RepeatedRRTForest(G.space_spec(g1["space"]), g1["start"], [0, 5], g1["goal"], g1["goal_radius"],
                           max_step=0.15, resolution=0.01, _backends=(OracleForest, OracleSpace))
    fn = str(tmp_path / "rrrt.csv")
    rows = testForest(rf, 700, fn)
    with open(fn) as f:
        lines = f.read().splitlines()
    assert lines[0] == "trial,plan iters,plan time,best cost"
    parsed = list(csv.reader(lines[1:]))
    assert all(len(r) == 4 for r in parsed) and len(parsed) == len(rows)
    assert [int(r[0]) for r in parsed] == sorted(int(r[0]) for r in parsed)          # grouped by trial
    t0 = [(int(r[1]), float(r[3])) for r in parsed if r[0] == "0"]
    want = [(c[0], c[1]) for c in g1["change_points"] if c[0] <= 700]
    assert t0[:-1] == want and t0[-1] == (700, want[-1][1])                           # closing row repeats the best
    assert [r for r in parsed if r[0] == "1"][-1][1] == "700"


def test_nearestneighbors_bookkeeping_without_reference():
    nn = fake_nn(P.MetricSpec.euclidean(2))
    assert nn.nearest([0, 0]) is None and nn.knearest([0, 0], 3) == [] and nn.neighbors([0, 0], 1) == []
    pts = [[0.1 * i, 0.05 * i] for i in range(20)]
    datas = [object() for _ in pts]
    for p, d in zip(pts, datas):
        nn.add(p, d)
    got = nn.nearest([0.52, 0.26])
    assert got[0] is pts[5] and got[1] is datas[5]           # the very objects that were stored
    assert nn.nearest([0.52, 0.26], lambda p, d: d is datas[5])[1] is datas[6 if abs(0.6 - 0.52) < abs(0.4 - 0.52) else 4]
    assert nn.remove(pts[5], datas[5]) == 1 and nn.remove(pts[5]) == 0
    assert nn.nearest([0.5, 0.25])[1] is not datas[5]
    assert [d for _, d in nn.knearest([0.0, 0.0], 3)] == datas[:3]
    assert [d for _, d in nn.neighbors([0.0, 0.0], 0.25)] == datas[:3]   # d <= radius, insertion order
    nn.inclusive_radius = False
    assert len(nn.neighbors([0.0, 0.0], nn.item(2) and (0.2 ** 2 + 0.1 ** 2) ** 0.5)) == 2
    everything = nn.knearest([0.0, 0.0], 500)
    assert len(everything) == 19
    nn.set(pts[:4])
    assert len(nn) == 4 and nn.nearest([10, 10])[0] is pts[3]
    nn.reset()
    assert nn.nearest([0, 0]) is None


def test_nearest_filter_falls_through_many_rejections():
    nn = fake_nn(P.MetricSpec.euclidean(1))
    for i in range(400):
        nn.add([float(i)], i)
    assert nn.nearest([0.0], lambda p, d: d < 300)[1] == 300   # > 128 rejected in a row -> device mask path
    assert nn.nearest([0.0], lambda p, d: True) is None
    assert [d for _, d in nn.knearest([0.0], 3, lambda p, d: d % 2 == 0)] == [1, 3, 5]


def test_compaction_keeps_answers():
    nn = fake_nn(P.MetricSpec.euclidean(2))
    rng = random.Random(3)
    pts = [[rng.random(), rng.random()] for _ in range(2500)]
    for i, p in enumerate(pts):
        nn.add(p, i)
    for i in range(0, 2500, 2):
        assert nn.remove(pts[i], i) == 1
    for i in range(1, 1500, 2):
        assert nn.remove(pts[i]) == 1       # triggers compaction (more than half dead)
    assert len(nn) == 500
    q = [0.3, 0.7]
    best = min((i for i in range(1501, 2500, 2)), key=lambda i: (pts[i][0] - q[0]) ** 2 + (pts[i][1] - q[1]) ** 2)
    assert nn.nearest(q)[1] == best


# ------------------------------------------------------------------ against the unmodified reference
@needs_ref
def test_descriptor_translation_of_reference_objects():
    pomp = refshim.load()
    from pomp.spaces import metric as M
    from pomp.planners import kinodynamicplanner as KP
    from pomp.example_problems import geometric, dubins, pendulum, doubleintegrator, flappy
    assert metric_from_pomp(M.euclideanMetric, 3).kind == _abi.METRIC_EUCLIDEAN
    assert metric_from_pomp(M.L1Metric, 2).kind == _abi.METRIC_L1
    assert metric_from_pomp(M.WeightedEuclideanMetric([1, 2]), 2).dim_weight == [1.0, 2.0]
    d = dubins.dubinsCarTest()
    m = metric_from_pomp(d.configurationSpace.distance)
    assert m.kind == _abi.METRIC_GEODESIC_PRODUCT and m.comp_kind == [0, 1] and abs(m.comp_weight[1] - 0.5 / np.pi) < 1e-18
    cm = KP.CostMetric(d.configurationSpace.distance, 0.0)
    cm.costMax = 2.5
    mc = metric_from_pomp(cm)
    assert mc.has_cost and mc.dim == 4 and mc.cost_max == 2.5 and mc.cost_weight == 0.0
    with pytest.raises(TypeError):
        metric_from_pomp(lambda a, b: 0.0, 2)
    s = space_from_pomp(geometric.kinkTest().configurationSpace)
    assert s.dim == 2 and len(s.boxes) == 4
    s = space_from_pomp(d.configurationSpace)
    assert s.dim == 3 and s.lo[2] == -np.inf
    s = space_from_pomp(pendulum.pendulumTest().configurationSpace)
    assert s.lo == [-np.inf, -10.0] and s.hi == [np.inf, 10.0]
    assert dynamics_from_pomp(d.controlSpace).kind == _abi.DYN_DUBINS
    assert dynamics_from_pomp(pendulum.pendulumTest().controlSpace).kind == _abi.DYN_PENDULUM
    di = dynamics_from_pomp(doubleintegrator.doubleIntegratorTest().controlSpace)
    assert di.kind == _abi.DYN_CV and di.x_dim == 4 and di.u_dim == 3 and di.dt == 0.05
    assert dynamics_from_pomp(flappy.flappyTest().controlSpace).kind == _abi.DYN_FLAPPY


@needs_ref
@pytest.mark.parametrize("D", [2, 3])
def test_dropin_nn_equals_reference_bruteforce(D):
    pomp = refshim.load()
    from pomp.spaces import metric as M
    from pomp.structures.nearestneighbors import NearestNeighbors as RefNN
    rng = random.Random(D)
    ref, mine = RefNN(M.euclideanMetric, "bruteforce"), fake_nn(M.euclideanMetric)
    refkd = RefNN(M.euclideanMetric, "kdtree")
    pts = [[rng.random() for _ in range(D)] for _ in range(400)]
    for i, p in enumerate(pts):
        ref.add(p, i)
        refkd.add(p, i)
        mine.add(p, i)
    for i in range(0, 400, 9):
        assert ref.remove(pts[i], i) == mine.remove(pts[i], i) == 1
        refkd.remove(pts[i], i)
    filt = lambda p, d: d % 5 == 0  # noqa: E731
    for _ in range(60):
        q = [rng.random() for _ in range(D)]
        assert ref.nearest(q) == mine.nearest(q)
        assert ref.nearest(q, filt) == mine.nearest(q, filt)
        assert refkd.knearest(q, 6) == mine.knearest(q, 6)
        assert refkd.knearest(q, 6, filt) == mine.knearest(q, 6, filt)
        assert sorted(refkd.neighbors(q, 0.2), key=lambda t: t[1]) == mine.neighbors(q, 0.2)


def _run_ref_planner(make, iters, install):
    import io
    import contextlib
    if install:
        P.install(fake_nn)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            planner = make()
            planner.planMore(iters)
        V, E = planner.getRoadmap()
        return roadmap_hash(V, E), len(V), planner
    finally:
        if install:
            P.uninstall()


@needs_ref
def test_unmodified_rrrt_planner_with_dropins_matches_golden():
    """r-rrt on Bugtrap (BASELINE config 1): the reference planner, driven through the drop-in
    NearestNeighbors + EpsilonEdgeChecker, reproduces the reference's own cost trace (G1)."""
    refshim.load()
    from pomp.example_problems import geometric
    import io
    import contextlib
    g = G.load("planners.json")["G1_rrrt_bugtrap_kdtree"]
    P.install(fake_nn)
    try:
        random.seed(0)
        prob = geometric.bugtrapTest()
        with contextlib.redirect_stdout(io.StringIO()):
            planner = prob.planner("r-rrt", nextStateSamplingRange=0.15, nearestNeighborMethod="b200",
                                   checker=fake_checker(prob.configurationSpace, 0.01))
            assert type(planner.nearestNeighbors).__name__ == "NearestNeighbors"
            assert planner.nearestNeighbors.method == "b200"
            cps, last, it = [], None, 0
            while it < 1000:
                planner.planMore(10)
                it += 10
                if planner.bestPathCost != last:
                    cps.append([it, planner.bestPathCost])
                    last = planner.bestPathCost
    finally:
        P.uninstall()
    assert cps == [c for c in g["change_points"] if c[0] <= 1000]


@needs_ref
def test_unmodified_sst_star_with_dropins_matches_reference():
    """sst* on DoubleIntegrator (BASELINE config 5): remove(), neighbors(), filters and the
    piecewise-linear edge checks all go through the drop-ins."""
    refshim.load()
    from pomp.example_problems import doubleintegrator

    def make(dropin):
        random.seed(7)
        prob = doubleintegrator.doubleIntegratorTest()
        kw = dict(selectionRadius=0.3, witnessRadius=0.3)
        if dropin:
            kw.update(nearestNeighborMethod="b200", checker=fake_checker(prob.configurationSpace, 0.01))
        return prob.planner("sst*", **kw)

    h_ref, n_ref, p_ref = _run_ref_planner(lambda: make(False), 1200, False)
    h_new, n_new, p_new = _run_ref_planner(lambda: make(True), 1200, True)
    assert (h_new, n_new) == (h_ref, n_ref)
    assert p_new.bestPathCost == p_ref.bestPathCost
    g = G.load("planners.json")["G3_sststar_doubleintegrator"]      # frozen output of the unmodified reference
    assert g["hash"] == h_new and g["num_nodes"] == n_new and g["best_cost"] == p_new.bestPathCost


@needs_ref
def test_unmodified_ao_rrt_dubins_with_dropins_matches_bruteforce_reference():
    """ao-rrt(numControlSamples=64) on Dubins (BASELINE config 4): CostMetric (mutated in place by
    the planner), DubinsCarInterpolator edges and the batched control selector.  The exact oracle
    for SO(2)/cost metrics is the reference's 'bruteforce' back end."""
    refshim.load()
    from pomp.example_problems import dubins

    def make(dropin):
        random.seed(7)
        prob = dubins.dubinsCarTest()
        if not dropin:
            return prob.planner("ao-rrt", numControlSamples=64, nearestNeighborMethod="bruteforce")
        planner = prob.planner("ao-rrt", numControlSamples=64, nearestNeighborMethod="b200",
                               checker=fake_checker(prob.configurationSpace, 0.01))
        sel = planner.rrt.controlSelector
        planner.setControlSelector(P.BatchedRandomControlSelector(
            sel.controlSpace, sel.metric, sel.numSamples, _backend_factory=lambda spec: OracleDynamics(spec)))
        return planner

    h_ref, n_ref, p_ref = _run_ref_planner(lambda: make(False), 400, False)
    h_new, n_new, p_new = _run_ref_planner(lambda: make(True), 400, True)
    assert (h_new, n_new) == (h_ref, n_ref)
    g = G.load("planners.json")["G4_aorrt64_dubins_bruteforce"]
    assert g["hash"] == h_new and g["num_nodes"] == n_new


@needs_ref
def test_unmodified_rrt_star_with_rewire_batch_matches_reference():
    """rrt* on Kink: knearest + the straight edges to / from every neighbour decided in one
    back-end call, then the reference's own rewire logic running on those cached answers -- same tree,
    same costs as the unmodified reference with its kd-tree, and far fewer edge-check calls."""
    refshim.load()
    from pomp.example_problems import geometric
    calls = {}

    def make(mode):
        random.seed(3)
        prob = geometric.kinkTest()
        if mode == "ref":
            return prob.planner("rrt*", nearestNeighborMethod="kdtree")
        chk = fake_checker(prob.configurationSpace, 0.01)
        planner = prob.planner("rrt*", nearestNeighborMethod="b200", checker=chk)
        be = chk._dev._be
        orig = be.edge_check_one
        calls[mode] = [0, chk]

        def counted(a, b, res, _o=orig, _c=calls[mode]):
            _c[0] += 1
            return _o(a, b, res)
        be.edge_check_one = counted
        if mode == "batch":
            planner.nearestNeighbors.enable_rewire_prefetch(chk)
        return planner

    class Chunked(object):   # planMore returns at every improvement (rrtstarplanner.py:86-90): keep going
        def __init__(self, p):
            self.p = p

        def planMore(self, iters):
            for _ in range(iters // 10):
                self.p.planMore(10)

        def __getattr__(self, k):
            return getattr(self.p, k)

    h_ref, n_ref, p_ref = _run_ref_planner(lambda: Chunked(make("ref")), 800, False)
    h_one, n_one, p_one = _run_ref_planner(lambda: Chunked(make("single")), 800, True)
    h_new, n_new, p_new = _run_ref_planner(lambda: Chunked(make("batch")), 800, True)
    assert (h_new, n_new) == (h_ref, n_ref) == (h_one, n_one) and n_ref > 100
    assert p_new.bestPathCost == p_ref.bestPathCost
    costs = lambda p: sorted(n.c for n in p.nodes) if hasattr(p, "nodes") else None
    assert costs(p_new) == costs(p_ref)
    assert calls["batch"][1].rewire_hits > 0
    assert calls["batch"][0] < calls["single"][0]      # the rewire edges no longer cost a call each


@needs_ref
def test_est_density_equals_reference_loop():
    """EST.density (kinodynamicplanner.py:688-693) through NearestNeighbors.density of the drop-in."""
    refshim.load()
    import math
    from pomp.structures.nearestneighbors import NearestNeighbors as RefNN
    from pomp.spaces.metric import euclideanMetric
    rng = random.Random(4)
    pts = [[rng.random(), rng.random()] for _ in range(1500)]
    ref = RefNN(euclideanMetric, "kdtree")
    mine = fake_nn(euclideanMetric)
    for i, p in enumerate(pts):
        ref.add(p, i)
        mine.add(p, i)
    radius = 0.03
    for _ in range(40):
        x = [rng.random(), rng.random()]
        want = sum(math.exp(-(euclideanMetric(nx, x) / radius) ** 2) for nx, n in ref.neighbors(x, radius * 4.0))
        got = mine.density(x, radius * 4.0, radius)
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want))


@needs_ref
def test_edge_checker_dispatch_on_reference_interpolators():
    refshim.load()
    from pomp.spaces.edgechecker import EpsilonEdgeChecker as RefChecker
    from pomp.spaces.interpolators import LambdaInterpolator
    from pomp.example_problems import geometric, dubins, pendulum
    rng = random.Random(5)
    kink = geometric.kinkTest().configurationSpace
    ref, mine = RefChecker(kink, 0.01), fake_checker(kink, 0.01)
    for _ in range(100):
        a, b = kink.sample(), kink.sample()
        e = kink.interpolator(a, b)
        assert ref.feasible(e) == mine.feasible(e) == mine.visible(a, b)
        lam = LambdaInterpolator(lambda u, a=a, b=b: [a[0] + u * (b[0] - a[0]), a[1] + u * u * (b[1] - a[1])], kink, 10)
        assert ref.feasible(lam) == mine.feasible(lam)
    d2 = dubins.dubinsTest2()
    ref, mine = RefChecker(d2.configurationSpace, 0.01), fake_checker(d2.configurationSpace, 0.01)
    for _ in range(100):
        x = d2.configurationSpace.sample()
        e = d2.controlSpace.interpolator(x, d2.controlSpace.controlSet(x).sample())
        assert ref.feasible(e) == mine.feasible(e)
    pp = pendulum.pendulumTest()
    ref, mine = RefChecker(pp.configurationSpace, 0.01), fake_checker(pp.configurationSpace, 0.01)
    for _ in range(60):
        x = [rng.uniform(0, 6.28), rng.uniform(-10.5, 10.5)]
        e = pp.controlSpace.interpolator(x, pp.controlSpace.controlSet(x).sample())
        assert ref.feasible(e) == mine.feasible(e)


@needs_ref
def test_sst_star_forest_trials_equal_the_reference_planner():
    """SSTStarForest (lock-step sst* trials, BASELINE config 5) with the oracle-backed stand-in for the device side:
    every trial reproduces the unmodified reference planner under its seed -- roadmap hash, node count, best cost --
    across the restart at iteration 300.  (The CUDA kernel is compared with the same stand-in on the GPU.)"""
    refshim.load()
    import io
    import contextlib
    from pomp.example_problems import doubleintegrator
    from pyoptimalmotionplanning_b200.sst import SSTStarForest
    from tests.fake_backend import OracleSSTForest
    seeds, iters = [7, 8, 3], 700
    want = []
    for s in seeds:
        random.seed(s)
        prob = doubleintegrator.doubleIntegratorTest()
        with contextlib.redirect_stdout(io.StringIO()):
            p = prob.planner("sst*", selectionRadius=0.3, witnessRadius=0.3)
            p.planMore(iters)
        V, E = p.getRoadmap()
        want.append((roadmap_hash(V, E), len(V), p.bestPathCost))
    prob = doubleintegrator.doubleIntegratorTest()
    state = random.getstate()
    f = SSTStarForest(prob.controlSpace, prob.start, prob.goal, seeds, prob.objective, selectionRadius=0.3,
                      witnessRadius=0.3, _backend_factory=OracleSSTForest)
    for _ in range(iters // 50):
        f.planMore(50)
    assert random.getstate() == state          # the global generator is untouched: every trial has its own stream
    for t in range(len(seeds)):
        V, E = f.getRoadmap(t)
        got = (hashlib.sha256(repr((V, E)).encode()).hexdigest()[:16], len(V), f.bestPathCost[t])
        assert got == want[t], (t, got, want[t])


@needs_ref
def test_projection_hash_density_equals_reference_est():
    """ESTWithProjections (kinodynamicplanner.py:746-868): the same bases, the same cells, the same densities."""
    refshim.load()
    import io
    import contextlib
    from pomp.planners import kinodynamicplanner as KP
    from pomp.example_problems import doubleintegrator, dubins
    from pomp.spaces.edgechecker import EpsilonEdgeChecker
    from tests.fake_backend import OracleProjectionDensity
    rng = random.Random(2)
    for prob in (doubleintegrator.doubleIntegratorTest(), dubins.dubinsCarTest()):
        cs = prob.controlSpace
        space = cs.configurationSpace()
        with contextlib.redirect_stdout(io.StringIO()):
            est = KP.ESTWithProjections(cs, EpsilonEdgeChecker(space, 0.01))
            est.setBoundaryConditions(prob.start, None)
        try:
            bounds = space.bounds()
        except Exception:
            bounds = None
        d = len(prob.start)
        mine = P.ProjectionHashDensity(d, bounds, _backend_factory=OracleProjectionDensity)
        mine.onAddNode(prob.start)
        xs = []
        for _ in range(300):
            x = [rng.uniform(-1.2, 1.2) for _ in range(d)]
            xs.append(x)
            with contextlib.redirect_stdout(io.StringIO()):
                est.onAddNode(KP.Node(x))
            mine.onAddNode(x)
        assert len(est.projectionBases) == len(mine.bases)
        for _ in range(100):
            x = rng.choice(xs) if rng.random() < 0.5 else [rng.uniform(-1.2, 1.2) for _ in range(d)]
            assert est.density(x) == mine.density(x)
