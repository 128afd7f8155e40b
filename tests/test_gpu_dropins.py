"""The Python drop-ins ON THE CUDA BACK END (no fake): structures.NearestNeighbors, spaces.EpsilonEdgeChecker,
DeviceControlSpace, BatchedRandomControlSelector and install() -- first against the oracle, then with the
UNMODIFIED reference planners driving them (r-rrt Bugtrap = BASELINE config 1, sst* DoubleIntegrator = config 5,
ao-rrt(64) Dubins = config 4, rrt* Kink), requiring the reference's own trees / costs.  The reference is the
verbatim copy staged by `make -C oracle ref` under oracle/_ref/ (see oracle/refshim.py).
"""
import contextlib
import hashlib
import io
import random

import numpy as np
import pytest

import pyoptimalmotionplanning_b200 as P
from oracle import oracle, refshim
from pyoptimalmotionplanning_b200 import MetricSpec, SpaceSpec, _abi
from tests import golden_util as G

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference copy not staged (make -C oracle ref)")


def roadmap_hash(V, E):
    return hashlib.sha256(repr((V, [(i, j) for i, j, u in E])).encode()).hexdigest()[:16]


# ------------------------------------------------------------------ NearestNeighbors vs the oracle
def test_nearestneighbors_dropin_on_cuda_matches_oracle():
    """add / remove / set / nearest(+filter) / knearest(+filter) / neighbors / compaction, brute-force regime and
    (after 20k points) the grid regime, against the oracle's masked scans."""
    rng = random.Random(1)
    m = MetricSpec.euclidean(2)
    nn = P.NearestNeighbors(m)
    assert nn.nearest([0, 0]) is None and nn.knearest([0, 0], 3) == [] and nn.neighbors([0, 0], 1) == []
    pts = [[rng.random(), rng.random()] for _ in range(3000)]
    datas = list(range(3000))
    for p, d in zip(pts, datas):
        nn.add(p, d)
    arr = np.asarray(pts)
    dead = np.zeros(3000, np.uint8)
    for i in range(0, 3000, 7):
        assert nn.remove(pts[i], i) == 1
        dead[i] = 1
    assert nn.remove(pts[0], 0) == 0
    filt = lambda p, d: d % 5 == 0  # noqa: E731
    fmask = dead.copy()
    fmask[::5] = 1
    for _ in range(40):
        q = [rng.random(), rng.random()]
        oi, _ = oracle.nn_nearest(m, arr, [q], dead)
        got = nn.nearest(q)
        assert got[0] is pts[oi[0]] and got[1] == oi[0]                 # the very objects that were stored
        oi, _ = oracle.nn_nearest(m, arr, [q], fmask)
        assert nn.nearest(q, filt)[1] == oi[0]
        ki, _, kc = oracle.nn_knearest(m, arr, [q], 6, dead)
        assert [d for _, d in nn.knearest(q, 6)] == list(ki[0, :kc[0]])
        ki, _, kc = oracle.nn_knearest(m, arr, [q], 6, fmask)
        assert [d for _, d in nn.knearest(q, 6, filt)] == list(ki[0, :kc[0]])
        off, hits = oracle.nn_neighbors(m, arr, [q], 0.05, True, dead)
        assert [d for _, d in nn.neighbors(q, 0.05)] == list(hits)
    # > 128 rejections in a row -> device mask path; everything rejected -> None
    assert nn.nearest([0.5, 0.5], lambda p, d: True) is None
    far = lambda p, d: (p[0] - 0.5) ** 2 + (p[1] - 0.5) ** 2 < 0.09  # noqa: E731
    fm = dead.copy()
    fm[[i for i, p in enumerate(pts) if far(p, None)]] = 1
    oi, _ = oracle.nn_nearest(m, arr, [[0.5, 0.5]], fm)
    assert nn.nearest([0.5, 0.5], far)[1] == oi[0]
    # compaction (more than half dead) renumbers internally but answers stay right; generation changes
    gen = nn.generation
    for i in range(3000):
        if not dead[i] and i % 3:
            assert nn.remove(pts[i]) == 1
            dead[i] = 1
    assert nn.generation > gen and len(nn) == int((dead == 0).sum())
    for _ in range(20):
        q = [rng.random(), rng.random()]
        oi, _ = oracle.nn_nearest(m, arr, [q], dead)
        assert nn.nearest(q)[1] == oi[0]
    # grid regime through the same object: bulk set + batched queries
    big = np.random.default_rng(2).random((60000, 2))
    nn.set(big.tolist(), list(range(60000)))
    qs = np.random.default_rng(3).random((500, 2))
    idx, dist = nn.nearest_batch(qs)
    oi, od = oracle.nn_nearest(m, big, qs)
    assert (idx == oi).all() and (dist == od).all()
    assert nn.nearest(qs[0].tolist())[1] == oi[0]
    kidx, kdist, kcnt = nn.knearest_batch(qs, 9)
    oki, okd, okc = oracle.nn_knearest(m, big, qs, 9)
    assert (kidx == oki).all() and (kdist == okd).all()
    off, hits = nn.neighbors_batch(qs, 0.01)
    ooff, ohits = oracle.nn_neighbors(m, big, qs, 0.01, True)
    assert (off == ooff).all() and (hits == ohits).all()
    nn.reset()
    assert nn.nearest([0, 0]) is None


def test_edge_checker_dropin_on_cuda_matches_oracle():
    rng = np.random.default_rng(5)
    boxes = np.concatenate([rng.random((300, 2)) * 0.9, np.zeros((300, 2))], 1)
    boxes[:, 2:] = boxes[:, :2] + 0.005 + rng.random((300, 2)) * 0.03
    circles = np.concatenate([rng.random((40, 2)), 0.01 + 0.02 * rng.random((40, 1))], 1)
    sp = SpaceSpec([0, 0], [1, 1], boxes=boxes, circles=circles)
    chk = P.EpsilonEdgeChecker(sp, 0.01)
    a = rng.random((4000, 2))
    b = np.clip(a + (rng.random((4000, 2)) - 0.5) * 0.3, -0.05, 1.05)
    want = oracle.edge_check(sp, a, b, 0.01)
    assert (chk.feasible_batch(a, b) == want).all()
    assert [chk.visible(a[i], b[i]) for i in range(200)] == [bool(w) for w in want[:200]]
    assert (chk._dev.feasible_batch(a) == oracle.feasible(sp, a)).all()
    assert [chk._dev.feasible(a[i]) for i in range(100)] == [bool(w) for w in oracle.feasible(sp, a[:100])]
    # piecewise-linear paths on the SE(2) geodesic and Dubins control edges
    g = MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi])
    sp3 = SpaceSpec([0, 0, -np.inf], [1, 1, np.inf], boxes=boxes)
    chk3 = P.EpsilonEdgeChecker(sp3, 0.01)
    wp = rng.random((600, 3)) * [1, 1, 6.28]
    off = np.arange(0, 601, 6)
    assert (chk3.feasible_paths(wp, off, g) == oracle.edge_check_path(sp3, g, wp, off, 0.01)).all()
    # a resolution that asks for > 1e7 samples is an error, not a hung GPU (reference: never returns)
    if type(chk._dev._be).__name__ == "CudaSpace":
        with pytest.raises(P.PompError):
            chk._dev._be.edge_check(a[:4], b[:4], 1e-12)
        assert (chk.feasible_batch(a[:50], b[:50]) == want[:50]).all()      # the flag was cleared


def test_control_space_and_selector_dropins_on_cuda_match_oracle():
    from pyoptimalmotionplanning_b200 import DynamicsSpec
    from pyoptimalmotionplanning_b200.backend import CudaDynamics
    rng = np.random.default_rng(6)
    dyn = DynamicsSpec(_abi.DYN_DUBINS, 3, 2, _abi.COST_ABS_U0)
    m = MetricSpec.geodesic_product([("R", 2), ("SO2",)], [1.0, 0.5 / np.pi]).with_cost(1.0, 3.0)
    x = rng.random((300, 4)) * [1, 1, 6.28, 1]
    xd = rng.random((300, 4)) * [1, 1, 6.28, 3]
    u = (rng.random((300, 64, 2)) - 0.5) * [0.2, 6.28]
    valid = (rng.random((300, 64)) < 0.9).astype(np.uint8)
    be = CudaDynamics(dyn)
    best, xb, db = be.select_control(m, x, xd, u, valid)
    ob, oxb, odb = oracle.select_control(dyn, m, x, xd, u, valid)
    assert (best == ob).all()
    np.testing.assert_allclose(xb, oxb, rtol=1e-5, atol=0, equal_nan=True)     # north-star state tolerance
    nx = be.next_state(x, u[:, 0])
    np.testing.assert_allclose(nx, oracle.next_state(dyn, x, u[:, 0]), rtol=1e-5, atol=1e-12)
    # LTI (controlspace.py:88-112): x' = A x + B u, numpy dot order
    A = rng.random((3, 3)) - 0.5
    B = rng.random((3, 2)) - 0.5
    lti = DynamicsSpec(_abi.DYN_LTI, 3, 2, A=A, B=B)
    xs, us = rng.random((500, 3)), rng.random((500, 2))
    got = CudaDynamics(lti).next_state(xs, us)
    np.testing.assert_allclose(got, oracle.next_state(lti, xs, us), rtol=1e-5, atol=1e-15)
    np.testing.assert_allclose(got, xs @ A.T + us @ B.T, rtol=1e-12, atol=1e-15)
    golds = [c for c in G.load("dynamics.json") if c["spec"]["kind"] == "lti"]   # LTIControlSpace.nextState outputs
    assert len(golds) == 2
    for gold in golds:
        got = CudaDynamics(G.dyn_spec(gold["spec"])).next_state(gold["x"], gold["u"])
        np.testing.assert_allclose(got, np.array(gold["xnext"]), rtol=1e-12, atol=1e-15)


# ------------------------------------------------------------------ the unmodified reference on the CUDA back end
def _run(make, iters, install):
    if install:
        P.install()
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
@pytest.mark.parametrize("D", [2, 3])
def test_dropin_nn_on_cuda_equals_reference_classes(D):
    refshim.load()
    from pomp.spaces import metric as M
    from pomp.structures.nearestneighbors import NearestNeighbors as RefNN
    rng = random.Random(D)
    ref, mine = RefNN(M.euclideanMetric, "bruteforce"), P.NearestNeighbors(M.euclideanMetric)
    refkd = RefNN(M.euclideanMetric, "kdtree")
    pts = [[rng.random() for _ in range(D)] for _ in range(400)]
    for i, p in enumerate(pts):
        ref.add(p, i)
        refkd.add(p, i)
        mine.add(p, i)
    with contextlib.redirect_stdout(io.StringIO()):
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


@needs_ref
def test_unmodified_rrrt_bugtrap_on_cuda_matches_golden():
    """BASELINE config 1: `main.py Bugtrap r-rrt` through install() + the CUDA drop-ins reproduces the cost trace the
    unmodified reference produced (tests/golden/planners.json, G1)."""
    refshim.load()
    from pomp.example_problems import geometric
    g = G.load("planners.json")["G1_rrrt_bugtrap_kdtree"]
    P.install()
    try:
        random.seed(0)
        prob = geometric.bugtrapTest()
        with contextlib.redirect_stdout(io.StringIO()):
            planner = prob.planner("r-rrt", nextStateSamplingRange=0.15, nearestNeighborMethod="b200",
                                   checker=P.make_checker(prob.configurationSpace, 0.01))
            assert type(planner.nearestNeighbors).__module__.startswith("pyoptimalmotionplanning_b200")
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
    from pyoptimalmotionplanning_b200.backend import launch_count
    assert launch_count() > 0


@needs_ref
def test_unmodified_sst_star_on_cuda_matches_reference():
    """BASELINE config 5 (one trial): remove(), neighbors(), filters, piecewise-linear edges on the GPU."""
    refshim.load()
    from pomp.example_problems import doubleintegrator

    def make(dropin):
        random.seed(7)
        prob = doubleintegrator.doubleIntegratorTest()
        kw = dict(selectionRadius=0.3, witnessRadius=0.3)
        if dropin:
            kw.update(nearestNeighborMethod="b200", checker=P.make_checker(prob.configurationSpace, 0.01))
        return prob.planner("sst*", **kw)

    h_ref, n_ref, p_ref = _run(lambda: make(False), 1200, False)
    h_new, n_new, p_new = _run(lambda: make(True), 1200, True)
    assert (h_new, n_new) == (h_ref, n_ref)
    assert p_new.bestPathCost == p_ref.bestPathCost
    g = G.load("planners.json")["G3_sststar_doubleintegrator"]      # frozen output of the unmodified reference
    assert g["hash"] == h_new and g["num_nodes"] == n_new and g["best_cost"] == p_new.bestPathCost


@needs_ref
def test_unmodified_ao_rrt_dubins_on_cuda_matches_bruteforce_reference():
    """BASELINE config 4: CostMetric mutated in place, DubinsCarInterpolator edges, the batched control selector."""
    refshim.load()
    from pomp.example_problems import dubins

    def make(dropin):
        random.seed(7)
        prob = dubins.dubinsCarTest()
        if not dropin:
            return prob.planner("ao-rrt", numControlSamples=64, nearestNeighborMethod="bruteforce")
        planner = prob.planner("ao-rrt", numControlSamples=64, nearestNeighborMethod="b200",
                               checker=P.make_checker(prob.configurationSpace, 0.01))
        sel = planner.rrt.controlSelector
        planner.setControlSelector(P.BatchedRandomControlSelector(sel.controlSpace, sel.metric, sel.numSamples))
        return planner

    h_ref, n_ref, p_ref = _run(lambda: make(False), 400, False)
    h_new, n_new, p_new = _run(lambda: make(True), 400, True)
    assert (h_new, n_new) == (h_ref, n_ref)
    g = G.load("planners.json")["G4_aorrt64_dubins_bruteforce"]
    assert g["hash"] == h_new and g["num_nodes"] == n_new


@needs_ref
def test_unmodified_rrt_star_on_cuda_with_rewire_batch_matches_reference():
    refshim.load()
    from pomp.example_problems import geometric

    def make(mode):
        random.seed(3)
        prob = geometric.kinkTest()
        if mode == "ref":
            return prob.planner("rrt*", nearestNeighborMethod="kdtree")
        chk = P.make_checker(prob.configurationSpace, 0.01)
        planner = prob.planner("rrt*", nearestNeighborMethod="b200", checker=chk)
        planner.nearestNeighbors.enable_rewire_prefetch(chk)
        make.chk = chk
        return planner

    class Chunked(object):
        def __init__(self, p):
            self.p = p

        def planMore(self, iters):
            for _ in range(iters // 10):
                self.p.planMore(10)

        def __getattr__(self, k):
            return getattr(self.p, k)

    h_ref, n_ref, p_ref = _run(lambda: Chunked(make("ref")), 800, False)
    h_new, n_new, p_new = _run(lambda: Chunked(make("batch")), 800, True)
    assert (h_new, n_new) == (h_ref, n_ref) and n_ref > 100
    assert p_new.bestPathCost == p_ref.bestPathCost
    assert make.chk.rewire_hits > 0


@needs_ref
def test_edge_checker_on_cuda_dispatches_reference_interpolators():
    refshim.load()
    from pomp.spaces.edgechecker import EpsilonEdgeChecker as RefChecker
    from pomp.spaces.interpolators import LambdaInterpolator
    from pomp.example_problems import geometric, dubins, pendulum
    rng = random.Random(5)
    kink = geometric.kinkTest().configurationSpace
    ref, mine = RefChecker(kink, 0.01), P.EpsilonEdgeChecker(kink, 0.01)
    for _ in range(100):
        a, b = kink.sample(), kink.sample()
        e = kink.interpolator(a, b)
        assert ref.feasible(e) == mine.feasible(e) == mine.visible(a, b)
        lam = LambdaInterpolator(lambda u, a=a, b=b: [a[0] + u * (b[0] - a[0]), a[1] + u * u * (b[1] - a[1])], kink, 10)
        assert ref.feasible(lam) == mine.feasible(lam)
    d2 = dubins.dubinsTest2()
    ref, mine = RefChecker(d2.configurationSpace, 0.01), P.EpsilonEdgeChecker(d2.configurationSpace, 0.01)
    for _ in range(100):
        x = d2.configurationSpace.sample()
        e = d2.controlSpace.interpolator(x, d2.controlSpace.controlSet(x).sample())
        assert ref.feasible(e) == mine.feasible(e)
    pp = pendulum.pendulumTest()
    ref, mine = RefChecker(pp.configurationSpace, 0.01), P.EpsilonEdgeChecker(pp.configurationSpace, 0.01)
    for _ in range(60):
        x = [rng.uniform(0, 6.28), rng.uniform(-10.5, 10.5)]
        e = pp.controlSpace.interpolator(x, pp.controlSpace.controlSet(x).sample())
        assert ref.feasible(e) == mine.feasible(e)
    # an obstacle added after construction reaches the device tables
    kink2 = geometric.kinkTest().configurationSpace
    mine2 = P.EpsilonEdgeChecker(kink2, 0.01)
    assert mine2.visible([0.05, 0.05], [0.05, 0.95])
    kink2.addObstacle(geometric.Box(0.0, 0.4, 0.1, 0.6))
    assert not mine2.visible([0.05, 0.05], [0.05, 0.95])
    assert RefChecker(kink2, 0.01).feasible(kink2.interpolator([0.05, 0.05], [0.05, 0.95])) is False


@needs_ref
def test_device_control_space_nextstate_matches_reference():
    refshim.load()
    from pomp.example_problems import dubins, doubleintegrator, pendulum
    rng = random.Random(9)
    for prob in (dubins.dubinsCarTest(), doubleintegrator.doubleIntegratorTest(), pendulum.pendulumTest()):
        cs = prob.controlSpace
        dev = P.DeviceControlSpace(cs)
        space = prob.configurationSpace
        for _ in range(50):
            x = space.sample()
            if not all(np.isfinite(x)):
                x = [rng.uniform(-3, 3) for _ in x]
            u = cs.controlSet(x).sample()
            np.testing.assert_allclose(dev.nextState(x, u), cs.nextState(x, u), rtol=1e-5, atol=1e-12)


@needs_ref
def test_sst_star_forest_on_cuda_reproduces_the_reference_per_seed():
    """BASELINE config 5: 16 seeded sst* trials on DoubleIntegrator in lock step on the GPU (csrc/sst.cu), driven by
    the caller's own problem objects.  Every trial must reproduce what the UNMODIFIED reference planner produced
    under its seed (frozen in tests/golden/planners.json: roadmap hash, node count, best path cost after 1000
    iterations, across the sst* restart at iteration 300), and the oracle-backed stand-in node for node."""
    refshim.load()
    from pomp.example_problems import doubleintegrator
    from pyoptimalmotionplanning_b200.sst import SSTStarForest
    from tests.fake_backend import OracleSSTForest
    g = G.load("planners.json")["G3_sststar_forest"]
    prob = doubleintegrator.doubleIntegratorTest()
    f = SSTStarForest(prob.controlSpace, prob.start, prob.goal, g["seeds"], prob.objective, selectionRadius=0.3,
                      witnessRadius=0.3)
    o = SSTStarForest(prob.controlSpace, prob.start, prob.goal, g["seeds"][:3], prob.objective, selectionRadius=0.3,
                      witnessRadius=0.3, _backend_factory=OracleSSTForest)
    done = 0
    for chunk in (7, 93, 200, 1, 399, 300):           # uneven chunks: the restart falls inside and between launches
        f.planMore(chunk)
        o.planMore(chunk)
        done += chunk
        for t in range(3):
            assert f._be.get_trial(t) == tuple(o._be.get_trial(t)), (done, t)
            assert f.bestPathCost[t] == o.bestPathCost[t]
    assert done == g["iters"]
    for t, want in enumerate(g["results"]):
        V, E = f.getRoadmap(t)
        got = [hashlib.sha256(repr((V, E)).encode()).hexdigest()[:16], len(V), f.bestPathCost[t]]
        assert got == want, (g["seeds"][t], got, want)


# ------------------------------------------------------------------ EST projection-hash density
def test_projection_hash_density_on_cuda_matches_oracle():
    """pomp_est_* (csrc/est.cu) vs the dictionary restatement of ESTWithProjections.density
    (kinodynamicplanner.py:817-856): identical floats, including negative coordinates (int() truncates towards zero),
    cell borders and the cost-space scale."""
    rng = np.random.default_rng(5)
    for dim, cost in ((2, False), (4, False), (5, True), (6, False)):
        lo = -rng.random(dim) - 0.5
        hi = rng.random(dim) + 0.5
        dens = P.ProjectionHashDensity(dim, (list(lo), list(hi)), cost_space=cost)
        pts = lo + rng.random((5000, dim)) * (hi - lo)
        pts[::17] = np.round(pts[::17] * 13) / 13                      # some nodes exactly on cell borders
        dens.add_batch(pts[:100])
        for p in pts[100:140]:
            dens.onAddNode(list(p))
        dens.add_batch(pts[140:])
        assert len(dens) == 5000
        q = np.concatenate([pts[::50], lo + rng.random((200, dim)) * (hi - lo)])
        got = dens.density_batch(q)
        want = oracle.est_projection_density(dens.bases, pts, q)
        assert (got == want).all()
        assert dens.density(list(q[3])) == want[3]
        dens.reset()
        assert len(dens) == 0 and dens.density(list(q[0])) == 0.0


@needs_ref
def test_projection_hash_density_on_cuda_equals_reference_est():
    """The unmodified ESTWithProjections next to the device density: same bases, same densities."""
    refshim.load()
    from pomp.planners import kinodynamicplanner as KP
    from pomp.example_problems import doubleintegrator
    from pomp.spaces.edgechecker import EpsilonEdgeChecker
    rng = random.Random(3)
    prob = doubleintegrator.doubleIntegratorTest()
    cs = prob.controlSpace
    space = cs.configurationSpace()
    with contextlib.redirect_stdout(io.StringIO()):
        est = KP.ESTWithProjections(cs, EpsilonEdgeChecker(space, 0.01))
        est.setBoundaryConditions(prob.start, None)
    d = len(prob.start)
    mine = P.ProjectionHashDensity(d, space.bounds())
    mine.onAddNode(prob.start)
    xs = [[rng.uniform(-1.2, 1.2) for _ in range(d)] for _ in range(400)]
    with contextlib.redirect_stdout(io.StringIO()):
        for x in xs:
            est.onAddNode(KP.Node(x))
    mine.add_batch(xs)
    for x in xs[:60] + [[rng.uniform(-1.2, 1.2) for _ in range(d)] for _ in range(60)]:
        assert est.density(x) == mine.density(x)
