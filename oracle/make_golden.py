"""Generates tests/golden/*.json by running the UNMODIFIED reference (/root/reference) in this
container.  TEST INFRASTRUCTURE ONLY.  Re-run with:  python oracle/make_golden.py

The reference ships no tests/golden vectors for the hot path, so these outputs of the reference
itself are what pins the C oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_golden.py).  Everything is seeded; floats are stored with repr() precision.
"""
import hashlib
import json
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

pomp = refshim.load()
from pomp.spaces import metric as M  # noqa: E402
from pomp.spaces.configurationspace import *  # noqa: E402,F401,F403
from pomp.spaces.so2space import SO2Space  # noqa: E402
from pomp.spaces.edgechecker import EpsilonEdgeChecker  # noqa: E402
from pomp.spaces.interpolators import LinearInterpolator, PiecewiseLinearInterpolator  # noqa: E402
from pomp.spaces.costspace import CostControlSpace  # noqa: E402
from pomp.spaces.controlspace import ControlSpaceAdaptor  # noqa: E402
from pomp.structures.nearestneighbors import NearestNeighbors  # noqa: E402
from pomp.planners import kinodynamicplanner as KP  # noqa: E402
from pomp.example_problems import geometric, dubins, pendulum, doubleintegrator, flappy  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dump(name, obj):
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", name, os.path.getsize(os.path.join(OUT, name)), "bytes")


def rnd_pt(rng, lo, hi):
    return [rng.uniform(a, b) for a, b in zip(lo, hi)]


# ------------------------------------------------------------------ metrics
def gen_metrics():
    rng = random.Random(101)
    cases = []

    def add(name, spec, fn, lo, hi, n=200, special=()):
        pairs = [(rnd_pt(rng, lo, hi), rnd_pt(rng, lo, hi)) for _ in range(n)] + list(special)
        cases.append({"name": name, "spec": spec, "a": [p[0] for p in pairs], "b": [p[1] for p in pairs],
                      "d": [fn(p[0], p[1]) for p in pairs]})

    for D in (2, 3, 6):
        add("euclid%d" % D, {"kind": "euclidean", "dim": D}, M.euclideanMetric, [0] * D, [1] * D)
    add("l1_3", {"kind": "l1", "dim": 3}, M.L1Metric, [-1] * 3, [1] * 3)
    add("linf_3", {"kind": "linf", "dim": 3}, M.LinfMetric, [-1] * 3, [1] * 3)
    w = [1.0, 0.25, 3.0]
    add("weighted_3", {"kind": "weighted", "w": w}, M.WeightedEuclideanMetric(w), [-1] * 3, [1] * 3)
    add("scaled_2", {"kind": "scaled", "dim": 2, "w": 0.3}, M.WeightedEuclideanMetric(0.3), [0] * 2, [1] * 2)
    # Dubins C-space: Geometric2D x SO2, weights [1, 0.5/pi]  (dubins.py:92-94)
    dspace = dubins.dubinsCarTest().configurationSpace
    sp = [([0.1, 0.2, -1e-20], [0.3, 0.4, 0.0]), ([0.1, 0.2, 0.0], [0.1, 0.2, math.pi * 2]),
          ([0.5, 0.5, -7.5], [0.5, 0.5, 11.0]), ([0.5, 0.5, math.pi], [0.5, 0.5, 0.0])]
    add("se2_dubins", {"kind": "product", "comps": [["R", 2], ["SO2"]], "w": [1, 0.5 / math.pi]},
        dspace.distance, [0, 0, -1], [1, 1, 8], special=sp)
    pspace = pendulum.pendulumTest().configurationSpace
    add("pendulum", {"kind": "product", "comps": [["SO2"], ["R", 1]], "w": [1, 1]},
        pspace.distance, [-7, -10], [7, 10])
    dispace = doubleintegrator.doubleIntegratorTest().configurationSpace
    add("cv_di", {"kind": "product", "comps": [["R", 2], ["R", 2]], "w": [1, 1]},
        dispace.distance, [0, 0, -1, -1], [1, 1, 1, 1])
    for cw, cm in ((0.0, None), (1.0, None), (1.0, 1.2), (None, 0.5)):
        cmx = KP.CostMetric(dspace.distance, cw)
        cmx.costMax = cm
        add("cost_se2_w%s_m%s" % (cw, cm),
            {"kind": "product", "comps": [["R", 2], ["SO2"]], "w": [1, 0.5 / math.pi],
             "cost": {"weight": cw, "max": cm}},
            cmx, [0, 0, 0, 0], [1, 1, 6.3, 2], n=120)
    cme = KP.CostMetric(M.euclideanMetric, 1.0)
    cme.costMax = 1.5
    add("cost_euclid2", {"kind": "euclidean", "dim": 2, "cost": {"weight": 1.0, "max": 1.5}}, cme,
        [0, 0, 0], [1, 1, 2], n=120)
    dump("metrics.json", cases)


# ------------------------------------------------------------------ nearest neighbours
def gen_nn():
    rng = random.Random(202)
    cases = []
    for name, D, N, metric, spec in (
            ("euclid2", 2, 300, M.euclideanMetric, {"kind": "euclidean", "dim": 2}),
            ("euclid3", 3, 257, M.euclideanMetric, {"kind": "euclidean", "dim": 3}),
            ("euclid6", 6, 200, M.euclideanMetric, {"kind": "euclidean", "dim": 6}),
            ("l1_2", 2, 150, M.L1Metric, {"kind": "l1", "dim": 2}),
    ):
        pts = [rnd_pt(rng, [0] * D, [1] * D) for _ in range(N)]
        # a few exact duplicates and lattice points to exercise ties
        pts[10] = list(pts[3])
        pts[20] = [0.5] * D
        pts[21] = [0.5] * D
        qs = [rnd_pt(rng, [-0.1] * D, [1.1] * D) for _ in range(40)] + [[0.5] * D, list(pts[3])]
        brute = NearestNeighbors(metric, "bruteforce")
        kd = NearestNeighbors(metric, "kdtree")
        ids = {}
        for i, p in enumerate(pts):
            brute.add(p, i)
            kd.add(p, i)
        rejected = set(rng.sample(range(N), N // 3))
        filt = lambda p, d: d in rejected  # noqa: E731
        radius = {2: 0.12, 3: 0.25, 6: 0.6}[D]
        c = {"name": name, "spec": spec, "pts": pts, "q": qs, "rejected": sorted(rejected), "radius": radius,
             "nearest_brute": [], "nearest_brute_filtered": [], "nearest_kd": [], "knearest_kd": [],
             "knearest_kd_filtered": [], "neighbors_brute_lt": [], "neighbors_kd_le": []}
        for q in qs:
            c["nearest_brute"].append(brute.nearest(q)[1])
            c["nearest_brute_filtered"].append(brute.nearest(q, filt)[1])
            c["nearest_kd"].append(kd.nearest(q)[1])
            c["knearest_kd"].append([d for (_, d) in kd.knearest(q, 7)])
            c["knearest_kd_filtered"].append([d for (_, d) in kd.knearest(q, 7, filt)])
            c["neighbors_brute_lt"].append([d for (_, d) in brute.neighbors(q, radius)])
            c["neighbors_kd_le"].append(sorted(d for (_, d) in kd.neighbors(q, radius)))
        cases.append(c)
    dump("nn.json", cases)


# ------------------------------------------------------------------ feasibility + edges
def space_spec(space):
    return {"lo": list(space.box.bmin), "hi": list(space.box.bmax),
            "boxes": [[o.bmin[0], o.bmin[1], o.bmax[0], o.bmax[1]] for o in space.obstacles
                      if isinstance(o, geometric.Box)],
            "circles": [[o.center[0], o.center[1], o.radius] for o in space.obstacles
                        if isinstance(o, geometric.Circle)]}


def gen_edges():
    rng = random.Random(303)
    cases = []
    spaces = {"kink": geometric.kinkTest().configurationSpace,
              "bugtrap": geometric.bugtrapTest().configurationSpace}
    circ = geometric.Geometric2DCSpace()
    circ.addObstacle(geometric.Circle(0.5, 0.4, 0.39))
    circ.addObstacle(geometric.Circle(0.2, 0.85, 0.1))
    circ.addObstacle(geometric.Box(0.7, 0.8, 0.9, 0.95))
    spaces["circles"] = circ
    for name, space in spaces.items():
        pts = [rnd_pt(rng, [-0.05, -0.05], [1.05, 1.05]) for _ in range(300)]
        # boundary-exact points
        for o in space.obstacles:
            if isinstance(o, geometric.Box):
                pts.append([o.bmin[0], o.bmin[1]])
                pts.append([o.bmax[0], 0.5 * (o.bmin[1] + o.bmax[1])])
        pts += [[0.0, 0.0], [1.0, 1.0], [1.0, 0.5]]
        feas = [bool(space.feasible(p)) for p in pts]
        edges = []
        for _ in range(250):
            a = rnd_pt(rng, [0, 0], [1, 1])
            if rng.random() < 0.6:  # RRT-style short edge
                b = rnd_pt(rng, [0, 0], [1, 1])
                d = math.hypot(b[0] - a[0], b[1] - a[1])
                if d > 0.15:
                    b = [a[0] + 0.15 / d * (b[0] - a[0]), a[1] + 0.15 / d * (b[1] - a[1])]
            else:
                b = rnd_pt(rng, [0, 0], [1, 1])
            edges.append((a, b))
        edges.append(([0.1, 0.1], [0.1, 0.1]))  # zero length
        out = {}
        for res in (0.01, 0.001, 0.05):
            chk = EpsilonEdgeChecker(space, res)
            out[str(res)] = [bool(chk.feasible(LinearInterpolator(a, b))) for a, b in edges]
        cases.append({"name": name, "space": space_spec(space), "pts": pts, "feasible": feas,
                      "a": [e[0] for e in edges], "b": [e[1] for e in edges], "edge_ok": out})
    dump("edges.json", cases)


# ------------------------------------------------------------------ dynamics
def gen_dynamics():
    rng = random.Random(404)
    random.seed(404)
    cases = []

    def run(name, spec, cs, xs, us):
        cases.append({"name": name, "spec": spec, "x": xs, "u": us,
                      "xnext": [cs.nextState(x, u) for x, u in zip(xs, us)]})

    # Dubins (dubins.py:107-125)
    dprob = dubins.dubinsCarTest()
    dcs = dprob.controlSpace
    xs = [rnd_pt(rng, [0, 0, -1], [1, 1, 7]) for _ in range(200)]
    us = [dcs.controlSet(x).sample() for x in xs]
    us[0] = [0.05, 0.0]
    us[1] = [0.05, 1e-9]
    us[2] = [-0.07, 2.0]
    us[3] = [0.07, -2.0]
    run("dubins", {"kind": "dubins"}, dcs, xs, us)
    dcost = CostControlSpace(dcs, dprob.objective)
    xc = [x + [rng.uniform(0, 2)] for x in xs]
    run("dubins_cost", {"kind": "dubins", "cost": "abs_u0"}, dcost, xc, us)
    # Pendulum (controlspace.py:249-267, pendulum.py:152-154)
    pprob = pendulum.pendulumTest()
    pcs = pprob.controlSpace
    xs = [rnd_pt(rng, [0, -10], [2 * math.pi, 10]) for _ in range(120)]
    us = [pcs.controlSet(x).sample() for x in xs]
    us[0] = [0.0, 2.0]
    us[1] = [0.005, -2.0]
    us[2] = [0.5, 2.0]
    run("pendulum", {"kind": "pendulum", "dt": pcs.dt, "p": [9.8, 1, 1]}, pcs, xs, us)
    pcost = CostControlSpace(pcs, pprob.objective)
    xc = [x + [rng.uniform(0, 2)] for x in xs]
    run("pendulum_cost", {"kind": "pendulum", "dt": pcs.dt, "p": [9.8, 1, 1], "cost": "time"}, pcost, xc, us)
    # Double integrator closed form + stepwise trajectory (statespace.py:47-70)
    iprob = doubleintegrator.doubleIntegratorTest()
    ics = iprob.controlSpace
    xs = [rnd_pt(rng, [0, 0, -1, -1], [1, 1, 1, 1]) for _ in range(150)]
    us = [ics.controlSet(x).sample() for x in xs]
    us[0] = [0.0, 1.0, -1.0]
    us[1] = [0.05, 5.0, -5.0]
    run("cv", {"kind": "cv", "x_dim": 4, "dt": ics.dt}, ics, xs, us)
    cases.append({"name": "cv_euler", "spec": {"kind": "cv_euler", "x_dim": 4, "dt": ics.dt}, "x": xs, "u": us,
                  "xnext": [ics.trajectory(x, u)[-1] for x, u in zip(xs, us)]})
    # Flappy (flappy.py:20-31)
    fprob = flappy.flappyTest()
    fcs = fprob.controlSpace
    xs = [rnd_pt(rng, [0, 0, -40], [1000, 600, 40]) for _ in range(100)]
    us = [fcs.controlSet(x).sample() for x in xs]
    run("flappy", {"kind": "flappy", "p": [5, -1, 4]}, fcs, xs, us)
    # Adaptor + path length cost (controlspace.py:78-79, objectives.py:25-26)
    gprob = geometric.kinkTest()
    acs = ControlSpaceAdaptor(gprob.configurationSpace)
    acost = CostControlSpace(acs, gprob.objective)
    xs = [rnd_pt(rng, [0, 0, 0], [1, 1, 3]) for _ in range(60)]
    us = [rnd_pt(rng, [0, 0], [1, 1]) for _ in range(60)]
    run("adaptor_cost", {"kind": "adaptor", "x_dim": 2, "cost": "path_length"}, acost, xs, us)
    # LTI (controlspace.py:88-112): the LQR example's double integrator (lqr.py:40-65) and a dense random system;
    # numpy's dot goes through BLAS (FMA / blocked order), so this case carries the north-star state tolerance
    import numpy as np
    from pomp.spaces.controlspace import LTIControlSpace
    from pomp.example_problems import lqr
    lprob = lqr.lqrTest()
    lcs = lprob.controlSpace
    xs = [rnd_pt(rng, [-2, -2], [2, 2]) for _ in range(80)]
    us = [[rng.uniform(-1, 1)] for _ in range(80)]
    cases.append({"name": "lti_lqr", "spec": {"kind": "lti", "A": lcs.A.tolist(), "B": lcs.B.tolist()}, "x": xs, "u": us,
                  "tol": 1e-12, "xnext": [lcs.nextState(x, u) for x, u in zip(xs, us)]})
    A = np.array([[rng.uniform(-1, 1) for _ in range(4)] for _ in range(4)])
    B = np.array([[rng.uniform(-1, 1) for _ in range(3)] for _ in range(4)])
    dense = LTIControlSpace(None, None, A, B)
    xs = [rnd_pt(rng, [-3] * 4, [3] * 4) for _ in range(80)]
    us = [rnd_pt(rng, [-2] * 3, [2] * 3) for _ in range(80)]
    cases.append({"name": "lti_dense", "spec": {"kind": "lti", "A": A.tolist(), "B": B.tolist()}, "x": xs, "u": us,
                  "tol": 1e-12, "xnext": [dense.nextState(x, u) for x, u in zip(xs, us)]})
    dump("dynamics.json", cases)


# ------------------------------------------------------------------ control edges (kinodynamic)
def gen_control_edges():
    rng = random.Random(505)
    random.seed(505)
    cases = []
    # Dubins2 has obstacles (dubins.py:158-180)
    d2 = dubins.dubinsTest2()
    dcs = d2.controlSpace
    chk = EpsilonEdgeChecker(d2.configurationSpace, 0.01)
    xs, us, ok = [], [], []
    for _ in range(300):
        x = rnd_pt(rng, [0, 0, 0], [1, 1, 2 * math.pi])
        u = dcs.controlSet(x).sample()
        xs.append(x)
        us.append(u)
        ok.append(bool(chk.feasible(dcs.interpolator(x, u))))
    cs2 = d2.configurationSpace.components[0]
    cases.append({"name": "dubins2", "dyn": {"kind": "dubins"},
                  "space": {"lo": list(cs2.box.bmin) + [None], "hi": list(cs2.box.bmax) + [None],
                            "boxes": space_spec(cs2)["boxes"], "circles": []},
                  "geodesic": {"kind": "product", "comps": [["R", 2], ["SO2"]], "w": [1, 0.5 / math.pi]},
                  "res": 0.01, "x": xs, "u": us, "ok": ok})
    # Pendulum: theta unbounded, omega in [-10,10]; PiecewiseLinearInterpolator w/ product geodesic
    pp = pendulum.pendulumTest()
    pcs = pp.controlSpace
    chk = EpsilonEdgeChecker(pp.configurationSpace, 0.01)
    xs, us, ok, paths = [], [], [], []
    for i in range(200):
        x = rnd_pt(rng, [0, -10.5], [2 * math.pi, 10.5])
        u = pcs.controlSet(x).sample()
        e = pcs.interpolator(x, u)
        xs.append(x)
        us.append(u)
        ok.append(bool(chk.feasible(e)))
        if i < 40:
            paths.append(e.path)
    cases.append({"name": "pendulum", "dyn": {"kind": "pendulum", "dt": pcs.dt, "p": [9.8, 1, 1]},
                  "space": {"lo": [None, -10.0], "hi": [None, 10.0], "boxes": [], "circles": []},
                  "geodesic": {"kind": "product", "comps": [["SO2"], ["R", 1]], "w": [1, 1]},
                  "res": 0.01, "x": xs, "u": us, "ok": ok, "paths": paths})
    # Double integrator on the Kink obstacle set (DoubleIntegrator has none of its own)
    kink = geometric.kinkTest().configurationSpace
    from pomp.spaces.statespace import CVControlSpace
    from pomp.spaces.sets import BoxSet
    ics = CVControlSpace(kink, BoxConfigurationSpace([-1, -1], [1, 1]), BoxSet([-5, -5], [5, 5]), dt=0.05, dtmax=0.5)
    chk = EpsilonEdgeChecker(ics.configurationSpace(), 0.01)
    xs, us, ok = [], [], []
    for _ in range(300):
        x = rnd_pt(rng, [0, 0, -1, -1], [1, 1, 1, 1])
        u = ics.controlSet(x).sample()
        xs.append(x)
        us.append(u)
        ok.append(bool(chk.feasible(ics.interpolator(x, u))))
    cases.append({"name": "cv_kink", "dyn": {"kind": "cv_euler", "x_dim": 4, "dt": 0.05},
                  "space": {"lo": [0, 0, -1, -1], "hi": [1, 1, 1, 1], "boxes": space_spec(kink)["boxes"],
                            "circles": []},
                  "geodesic": {"kind": "product", "comps": [["R", 2], ["R", 2]], "w": [1, 1]},
                  "res": 0.01, "x": xs, "u": us, "ok": ok})
    dump("control_edges.json", cases)


# ------------------------------------------------------------------ control selection
def gen_select():
    cases = []
    for name, prob, spec, S in (
            ("dubins", dubins.dubinsCarTest(), {"kind": "dubins", "cost": "abs_u0"}, 64),
            ("pendulum", pendulum.pendulumTest(), {"kind": "pendulum", "dt": 0.01, "p": [9.8, 1, 1], "cost": "time"}, 64),
            ("cv", doubleintegrator.doubleIntegratorTest(), {"kind": "cv", "x_dim": 4, "dt": 0.05, "cost": "time"}, 16),
    ):
        cs = CostControlSpace(prob.controlSpace, prob.objective)
        base = M.euclideanMetric if prob.euclidean else prob.configurationSpace.distance
        metric = KP.CostMetric(base, 1.0)
        metric.costMax = 3.0
        sel = KP.RandomControlSelector(cs, metric, S)
        rng = random.Random(606)
        rows = []
        for t in range(40):
            x = prob.configurationSpace.sample() + [rng.uniform(0, 2.5)]
            xd = prob.configurationSpace.sample() + [rng.uniform(0, 3.5)]
            random.seed(1000 + t)
            ubest = sel.select(x, xd)
            random.seed(1000 + t)
            U = cs.controlSet(x)
            cands = [U.sample() for _ in range(S)]
            valid = [bool(U.contains(u)) for u in cands]
            best = -1 if ubest is None else cands.index(ubest)
            rows.append({"x": x, "xdes": xd, "u": cands, "valid": valid, "best": best})
        if name == "dubins":
            mspec = {"kind": "product", "comps": [["R", 2], ["SO2"]], "w": [1, 0.5 / math.pi]}
        elif name == "pendulum":
            mspec = {"kind": "product", "comps": [["SO2"], ["R", 1]], "w": [1, 1]}
        else:
            mspec = {"kind": "euclidean", "dim": 4}
        mspec["cost"] = {"weight": 1.0, "max": 3.0}
        cases.append({"name": name, "dyn": spec, "metric": mspec, "S": S, "rows": rows})
    dump("select_control.json", cases)


# ------------------------------------------------------------------ planner goldens (G1, G2)
def roadmap_hash(V, E):
    return hashlib.sha256(repr((V, [(i, j) for i, j, u in E])).encode()).hexdigest()[:16]


def gen_planners():
    out = {}
    # G2: plain RRT exploration trees
    for pname, prob in (("bugtrap", geometric.bugtrapTest()), ("kink", geometric.kinkTest())):
        for method in ("bruteforce", "kdtree"):
            random.seed(12345)
            space = prob.configurationSpace
            cs = ControlSpaceAdaptor(space)
            cs.nextStateSamplingRange = 0.15
            rrt = KP.RRT(cs, M.euclideanMetric, EpsilonEdgeChecker(space, 0.01), nearestNeighborMethod=method)
            rrt.setControlSelector(KP.KinematicControlSelector(cs, 0.15))
            rrt.setBoundaryConditions(prob.start, None)
            rrt.planMore(2000)
            V, E = rrt.getRoadmap()
            parents = [None] * len(rrt.nodes)
            index = {id(n): i for i, n in enumerate(rrt.nodes)}
            for i, n in enumerate(rrt.nodes):
                parents[i] = -1 if n.parent is None else index[id(n.parent)]
            key = "G2_%s_%s" % (pname, method)
            out[key] = {"seed": 12345, "iters": 2000, "n": len(V), "hash": roadmap_hash(V, E),
                        "start": prob.start, "space": space_spec(space)}
            if method == "bruteforce":
                # nodes in insertion order + parent indices: enough to rebuild getRoadmap()
                out[key]["nodes"] = [n.x for n in rrt.nodes]
                out[key]["parents"] = parents
    # G1: r-rrt on Bugtrap, change points of (iters, bestPathCost)
    for method in ("bruteforce", "kdtree"):
        random.seed(0)
        prob = geometric.bugtrapTest()
        planner = prob.planner("r-rrt", nextStateSamplingRange=0.15, nearestNeighborMethod=method)
        cps, last, it = [], None, 0
        for _ in range(300):
            planner.planMore(10)
            it += 10
            c = planner.bestPathCost
            if c != last:
                cps.append([it, c])
                last = c
        out["G1_rrrt_bugtrap_%s" % method] = {"seed": 0, "iters": it, "change_points": cps,
                                               "start": prob.start, "goal": prob.goal.c, "goal_radius": prob.goal.r,
                                               "space": space_spec(prob.configurationSpace)}
    # G3: sst* on DoubleIntegrator (BASELINE config 5), G4: ao-rrt(64) on Dubins with the exact (brute-force) NN
    # (BASELINE config 4): roadmap hash, size and best cost under a fixed seed
    import io
    import contextlib
    for seed in (7, 8):
        random.seed(seed)
        prob = doubleintegrator.doubleIntegratorTest()
        with contextlib.redirect_stdout(io.StringIO()):
            planner = prob.planner("sst*", selectionRadius=0.3, witnessRadius=0.3)
            planner.planMore(1200)
        V, E = planner.getRoadmap()
        key = "G3_sststar_doubleintegrator" + ("" if seed == 7 else "_seed%d" % seed)
        out[key] = {"seed": seed, "iters": 1200, "num_nodes": len(V), "hash": roadmap_hash(V, E),
                    "best_cost": planner.bestPathCost}
    # G3 forest: 16 seeded sst* trials (BASELINE config 5, "4096 independent seeded trials"), one reference run each
    seeds = list(range(100, 116))
    rows = []
    for seed in seeds:
        random.seed(seed)
        prob = doubleintegrator.doubleIntegratorTest()
        with contextlib.redirect_stdout(io.StringIO()):
            planner = prob.planner("sst*", selectionRadius=0.3, witnessRadius=0.3)
            planner.planMore(1000)
        V, E = planner.getRoadmap()
        rows.append([roadmap_hash(V, E), len(V), planner.bestPathCost])
    out["G3_sststar_forest"] = {"seeds": seeds, "iters": 1000, "results": rows}
    random.seed(7)
    prob = dubins.dubinsCarTest()
    with contextlib.redirect_stdout(io.StringIO()):
        planner = prob.planner("ao-rrt", numControlSamples=64, nearestNeighborMethod="bruteforce")
        planner.planMore(400)
    V, E = planner.getRoadmap()
    out["G4_aorrt64_dubins_bruteforce"] = {"seed": 7, "iters": 400, "num_nodes": len(V), "hash": roadmap_hash(V, E),
                                           "best_cost": planner.bestPathCost}
    dump("planners.json", out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    import io
    import contextlib
    gen_metrics()
    gen_nn()
    gen_edges()
    gen_dynamics()
    gen_control_edges()
    gen_select()
    with contextlib.redirect_stdout(io.StringIO()):
        pass
    gen_planners()
