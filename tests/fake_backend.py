"""Oracle-backed stand-ins for backend.CudaNN / CudaSpace / CudaDynamics -- TEST ONLY.

They implement the same small internal interface with the CPU oracle so that the HOST logic of
the drop-ins (index bookkeeping, filters, metric refresh, interpolator dispatch, RNG order, the
RRT driver, the multi-rank plumbing) can be exercised in the CPU-only container, including against
the unmodified reference planners.  The product never imports this module."""
import math

import numpy as np

from oracle import oracle


class OracleNN(object):
    def __init__(self, spec, device=0, capacity_hint=0):
        self.spec = spec
        self.dim = spec.dim
        self.pts = np.zeros((0, spec.dim))
        self.dead = np.zeros(0, np.uint8)
        self.rej = None

    def _mask(self):
        m = self.dead.copy()
        if self.rej is not None:
            m |= self.rej
        return m if m.any() else None

    def reset(self):
        self.pts = np.zeros((0, self.dim))
        self.dead = np.zeros(0, np.uint8)
        self.rej = None

    def size(self):
        return len(self.pts)

    def add(self, pts):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, self.dim)
        first = len(self.pts)
        self.pts = np.concatenate([self.pts, pts])
        self.dead = np.concatenate([self.dead, np.zeros(len(pts), np.uint8)])
        return first

    def add_one(self, pt):
        return self.add([list(pt)])

    def set(self, pts):
        self.reset()
        self.add(pts)

    def remove(self, index):
        self.dead[index] = 1

    def set_mask(self, mask):
        self.rej = None if mask is None else np.asarray(mask, dtype=np.uint8).copy()
        if self.rej is not None:
            assert len(self.rej) == len(self.pts)

    def update_metric(self, spec):
        self.spec = spec

    def nearest_one(self, pt):
        if len(self.pts) == 0:
            return -1, float("inf")
        i, d = oracle.nn_nearest(self.spec, self.pts, [list(pt)], self._mask())
        return int(i[0]), float(d[0])

    def nearest(self, q):
        return oracle.nn_nearest(self.spec, self.pts, q, self._mask())

    def knearest(self, q, k):
        return oracle.nn_knearest(self.spec, self.pts, q, k, self._mask())

    def neighbors(self, q, radius, inclusive=True):
        return oracle.nn_neighbors(self.spec, self.pts, q, radius, inclusive, self._mask())

    def density(self, q, radius, sigma, inclusive=True):
        return oracle.nn_density(self.spec, self.pts, q, radius, sigma, inclusive, self._mask())

    def rewire_candidates(self, space, x, k, res):
        ids, dist, cnt = oracle.nn_knearest(self.spec, self.pts, [list(x)], k, self._mask())
        c = int(cnt[0])
        ids, dist = ids[0, :c], dist[0, :c]
        nb = self.pts[ids]
        xs = np.repeat(np.asarray([list(x)], dtype=np.float64), c, 0)
        ok_in = oracle.edge_check(space.spec, nb, xs, res) if c else np.zeros(0, bool)
        ok_out = oracle.edge_check(space.spec, xs, nb, res) if c else np.zeros(0, bool)
        return ids, dist, np.asarray(ok_in, bool), np.asarray(ok_out, bool)


class OracleSpace(object):
    def __init__(self, spec, device=0):
        self.spec = spec
        self.dim = spec.dim

    def set_bounds(self, lo, hi):
        self.spec.lo, self.spec.hi = list(lo), list(hi)

    def feasible(self, x):
        return oracle.feasible(self.spec, x)

    def feasible_one(self, x):
        return bool(oracle.feasible(self.spec, [list(x)])[0])

    def edge_check(self, a, b, res):
        return oracle.edge_check(self.spec, a, b, res)

    def edge_check_one(self, a, b, res):
        return bool(oracle.edge_check(self.spec, [list(a)], [list(b)], res)[0])

    def edge_check_path(self, g, wp, off, res):
        return oracle.edge_check_path(self.spec, g, wp, off, res)

    def edge_check_control(self, dyn, g, x, u, res):
        if g is None:
            g = oracle.MetricSpec.euclidean(self.dim)
        return oracle.edge_check_control(self.spec, dyn, g, x, u, res)


class OracleDynamics(object):
    def __init__(self, spec):
        self.spec = spec

    def next_state(self, x, u):
        return oracle.next_state(self.spec, x, u)

    def select_control(self, m, x, xdes, u, valid=None):
        return oracle.select_control(self.spec, m, x, xdes, u, valid)


def oracle_rrt_extend(nn, space, xrand, max_step, res, check_sample=True):
    if check_sample and not space.feasible_one(xrand):
        return -2, [float("nan")] * nn.dim, 0.0, False
    parent, xnew, added = oracle.rrt_extend(nn.spec, nn.pts, space.spec, xrand, max_step, res, nn._mask())
    if parent < 0:
        return -1, [float("nan")] * nn.dim, 0.0, False
    a = nn.pts[parent]
    s = 0.0
    for j in range(nn.dim):
        s = s + (a[j] - xnew[j]) * (a[j] - xnew[j])
    elen = float(np.sqrt(s))
    xnew = [float(v) for v in xnew]
    if added:
        nn.add([xnew])
    return parent, xnew, elen, added


class OracleForest(object):
    """CPU stand-in for backend.CudaForest (same buffers, one oracle extend per trial)."""

    def __init__(self, dim, n_trials, capacity=16384, device=0):
        self.dim, self.T = dim, n_trials
        T, D = n_trials, dim
        self.xrand = [0.0] * (T * D)
        self.root = [0.0] * D
        self.active = [1] * T
        self.reset = [0] * T
        self.parent = [-1] * T
        self.xnew = [0.0] * (T * D)
        self.elen = [0.0] * T
        self.added = [0] * T
        self._nn = [None] * T

    def extend(self, space, max_step, resolution, check_sample=True):
        D = self.dim
        for t in range(self.T):
            if self.reset[t] or self._nn[t] is None:
                self._nn[t] = OracleNN(oracle.MetricSpec.euclidean(D))
                self._nn[t].add([list(self.root)])
            if not self.active[t]:
                self.parent[t], self.added[t] = -3, 0
                continue
            xr = self.xrand[t * D:(t + 1) * D]
            par, xn, el, ad = oracle_rrt_extend(self._nn[t], space, xr, max_step, resolution, check_sample)
            self.parent[t], self.elen[t], self.added[t] = par, el, 1 if ad else 0
            self.xnew[t * D:(t + 1) * D] = xn

    MAX_MULTI = 32

    def extend_multi(self, space, n_iters, max_step, resolution, check_sample=True, goal=None, goal_radius=0.0):
        import math
        T, D, M = self.T, self.dim, int(n_iters)
        self.n_done = [0] * T
        self.parent_m = [-1] * (T * self.MAX_MULTI)
        self.xnew_m = [0.0] * (T * self.MAX_MULTI * D)
        self.elen_m = [0.0] * (T * self.MAX_MULTI)
        self.added_m = [0] * (T * self.MAX_MULTI)
        for t in range(T):
            if not self.active[t]:
                continue
            if self.reset[t] or self._nn[t] is None:
                self._nn[t] = OracleNN(oracle.MetricSpec.euclidean(D))
                self._nn[t].add([list(self.root)])
            for it in range(M):
                xr = self.xrand_m[(t * M + it) * D:(t * M + it + 1) * D]
                par, xn, el, ad = oracle_rrt_extend(self._nn[t], space, xr, max_step, resolution, check_sample)
                o = t * M + it
                self.parent_m[o], self.elen_m[o], self.added_m[o] = par, el, 1 if ad else 0
                self.xnew_m[o * D:(o + 1) * D] = xn
                self.n_done[t] = it + 1
                if ad and goal is not None:
                    s = 0.0
                    for a, b in zip(xn, goal):
                        s = s + (a - b) * (a - b)
                    if math.sqrt(s) <= goal_radius:
                        break

    def _alloc_multi(self):
        if not hasattr(self, "xrand_m"):
            self.xrand_m = [0.0] * (self.T * self.MAX_MULTI * self.dim)


class OracleSSTForest(object):
    """CPU stand-in for backend.CudaSSTForest (csrc/sst.cu): the per-trial sst* iteration restated in Python on top of
    the oracle's pinned primitives (metric, stepwise CV trajectory end state, sampled control-edge check).  TEST ONLY:
    it validates the division of labour of sst.SSTStarForest against the unmodified reference where no GPU exists,
    and it is what the CUDA kernel is compared with on the GPU."""

    def __init__(self, x_dim, u_dim, n_trials, node_cap, witness_cap, dyn, geo, space_spec, device=0):
        self.X, self.U, self.T = x_dim, u_dim, n_trials
        self.dyn, self.geo, self.space = dyn, geo, space_spec
        self.euclid = oracle.MetricSpec.euclidean(x_dim)
        self.best = [float("inf")] * n_trials
        self.trials = None

    def _fresh(self, root):
        return {"x": [list(root)], "parent": [-1], "alive": [True], "active": [True], "c": [0.0], "nch": [0],
                "wx": [list(root)], "wrep": [0]}

    def reset(self, root):
        self.root = list(root)
        self.trials = [self._fresh(root) for _ in range(self.T)]
        self.best = [float("inf")] * self.T

    @staticmethod
    def _dist(a, b):
        s = 0.0
        for i in range(len(a)):
            s = s + (a[i] - b[i]) * (a[i] - b[i])
        return math.sqrt(s)

    def _iterate(self, tr, t, xrand, u, valid, rsel, rwit, goal_c, goal_r, res):
        X = self.X
        # pickNode (rrtstarplanner.py:390-410)
        best_i, best_c = -1, float("inf")
        for i in range(len(tr["x"])):
            if tr["alive"][i] and tr["active"][i] and self._dist(tr["x"][i], xrand) <= rsel and tr["c"][i] < best_c:
                best_i, best_c = i, tr["c"][i]
        if best_i < 0:
            bd = float("inf")
            for i in range(len(tr["x"])):
                if tr["alive"][i] and tr["active"][i]:
                    d = self._dist(tr["x"][i], xrand)
                    if d < bd:
                        best_i, bd = i, d
        if best_i < 0:
            return 1
        if not valid:
            return 0
        xn = tr["x"][best_i]
        if not oracle.edge_check_control(self.space, self.dyn, self.geo, [xn], [u], res)[0]:
            return 0
        xnew = oracle.next_state(self.dyn, [xn], [u])[0].tolist()
        newcost = tr["c"][best_i] + u[0]
        # nodeLocallyBest (:345-366)
        wi, wd = -1, float("inf")
        for j in range(len(tr["wx"])):
            d = self._dist(tr["wx"][j], xnew)
            if d < wd:
                wi, wd = j, d
        if wd > rwit:
            tr["wx"].append(list(xnew))
            tr["wrep"].append(-1)
            wi = len(tr["wx"]) - 1
        else:
            better = False
            for j in range(len(tr["wx"])):
                if self._dist(tr["wx"][j], xnew) <= rwit:
                    r = tr["wrep"][j]
                    if r < 0 or newcost < tr["c"][r]:
                        better = True
            if not better:
                return 0
        n = len(tr["x"])
        tr["x"].append(xnew)
        tr["parent"].append(best_i)
        tr["alive"].append(True)
        tr["active"].append(True)
        tr["c"].append(newcost)
        tr["nch"].append(0)
        tr["nch"][best_i] += 1
        # doPruning (:368-388): the NEAREST witness gets the new representative
        npeer = tr["wrep"][wi]
        tr["wrep"][wi] = n
        if npeer >= 0 and npeer != n:
            tr["active"][npeer] = False
            while npeer >= 0 and tr["nch"][npeer] == 0 and not tr["active"][npeer]:
                p = tr["parent"][npeer]
                tr["alive"][npeer] = False
                if p >= 0:
                    tr["nch"][p] -= 1
                tr["parent"][npeer] = -1
                npeer = p
        if goal_r >= 0 and oracle.metric(self.geo, xnew, goal_c) <= goal_r:
            if newcost < self.best[t]:
                self.best[t] = newcost
        return 0

    def step(self, M, xr, uu, valid, sel, wit, restart, goal_c, goal_r, res, restart_now=False, half_sq=None):
        none = [0] * self.T
        for t in range(self.T):
            if restart_now:
                self.trials[t] = self._fresh(self.root)
            for m in range(M):
                if restart[m]:
                    self.trials[t] = self._fresh(self.root)
                none[t] += self._iterate(self.trials[t], t, [float(v) for v in xr[t][m]], [float(v) for v in uu[t][m]],
                                         bool(valid[t][m]), sel[m], wit[m], goal_c, goal_r, res)
        nodes = [sum(1 for a in tr["alive"] if a) for tr in self.trials]
        return list(self.best), nodes, none

    def get_trial(self, trial):
        tr = self.trials[trial]
        return tr["x"], tr["parent"], tr["alive"]


class OracleProjectionDensity(object):
    """Stand-in for backend.CudaProjectionDensity.  TEST ONLY."""

    def __init__(self, dim, bases, device=0):
        self.dim, self.bases, self.pts = dim, bases, []

    def reset(self):
        self.pts = []

    def size(self):
        return len(self.pts)

    def add(self, pts):
        self.pts += [list(p) for p in np.asarray(pts, dtype=np.float64).reshape(-1, self.dim)]

    def density(self, q):
        return oracle.est_projection_density(self.bases, self.pts, np.asarray(q, dtype=np.float64).reshape(-1, self.dim))
