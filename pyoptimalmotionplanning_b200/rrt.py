"""Kinematic RRT / repeated-RRT driver on the fused extend kernel.

This is the host loop around `pomp_rrt_extend_h` (one launch + one sync per iteration).  It
mirrors the control flow of the reference callers for geometric (box-bounded, Euclidean)
problems so that, under the same `random.seed`, it grows THE SAME TREE:

  RRT.expand            pomp/planners/kinodynamicplanner.py:354-411
  RRT.planMore          :343-352
  RepeatedRRT.planMore  :1291-1326   ("r-rrt": restart the tree whenever the goal is reached)
  TreePlanner.getRoadmap :211-238

All randomness is drawn on the host from Python's global `random` in the reference's order:
random.uniform(0,1) for the goal bias (only when a goal is set, :358), then either the goal
neighbourhood (NeighborhoodSubset.sample -> sampleNeighborhood: x_i + uniform(-r,r),
configurationspace.py:38-44) or the box (BoxSet.sample: uniform(a_i,b_i), sets.py:175-176).
"""
import hashlib
import math
import collections
import random

from .descriptors import MetricSpec, space_from_pomp


class RRT(object):
    def __init__(self, space, start, goal=None, goal_radius=None, max_step=0.1, resolution=0.01, p_choose_goal=0.1,
                 device=0, _backends=None):
        """space: SpaceSpec (or pomp ConfigurationSpace); goal: centre of a NeighborhoodSubset of
        radius goal_radius, or None to just explore."""
        self.spec = space_from_pomp(space)
        self.D = self.spec.dim
        self.start = list(start)
        self.goal = None if goal is None else list(goal)
        self.goal_radius = goal_radius
        self.max_step = max_step
        self.resolution = resolution
        self.pChooseGoal = p_choose_goal
        if _backends is None:  # the product path: CUDA kernels behind the C-ABI
            from .backend import CudaNN, CudaSpace, rrt_extend
            _backends = (CudaNN, CudaSpace, rrt_extend)
        nn_cls, space_cls, self._extend = _backends
        self._nn = nn_cls(MetricSpec.euclidean(self.D), device)
        self._space = space_cls(self.spec, device)
        self.numIters = 0
        self._reset_tree()

    # ---- tree storage: parallel arrays in insertion order (= NearestNeighbors indices)
    def _reset_tree(self):
        self.nodes = [self.start]
        self.parents = [-1]
        self.costs = [0.0]
        self._nn.reset()
        self._nn.add_one(self.start)

    def reset(self):
        self._reset_tree()
        self.numIters = 0

    def _sample(self):
        if self.goal is not None and random.uniform(0.0, 1.0) < self.pChooseGoal:
            r = self.goal_radius
            return [c + random.uniform(-r, r) for c in self.goal]
        return [random.uniform(a, b) for a, b in zip(self.spec.lo, self.spec.hi)]

    def expand(self):
        """One RRT.expand; returns the new node index or None."""
        xrand = self._sample()
        parent, xnew, elen, added = self._extend(self._nn, self._space, xrand, self.max_step, self.resolution, True)
        if not added:
            return None
        self.nodes.append(xnew)
        self.parents.append(parent)
        self.costs.append(self.costs[parent] + elen)  # PathLengthObjectiveFunction.incremental
        return len(self.nodes) - 1

    def goal_contains(self, x):
        # NeighborhoodSubset.contains: space.distance(x, c) <= r (configurationspace.py:263-264)
        s = 0
        for a, b in zip(x, self.goal):
            s = s + (a - b) * (a - b)
        return math.sqrt(s) <= self.goal_radius

    def planMore(self, iters):
        for _ in range(iters):
            self.numIters += 1
            n = self.expand()
            if n is not None and self.goal is not None and self.goal_contains(self.nodes[n]):
                return True
        return False

    def getRoadmap(self):
        """(V, E) in the reference's traversal order (kinodynamicplanner.py:211-238)."""
        children = [[] for _ in self.nodes]
        for i, p in enumerate(self.parents):
            if p >= 0:
                children[p].append(i)
        V = [self.nodes[0]]
        E = []
        stack = [(0, 0)]
        while stack:
            n, i = stack.pop()
            for c in children[n]:
                j = len(V)
                E.append((i, j, self.nodes[c]))
                V.append(self.nodes[c])
                stack.append((c, j))
        return V, E

    def roadmap_hash(self):
        V, E = self.getRoadmap()
        return hashlib.sha256(repr((V, [(i, j) for i, j, u in E])).encode()).hexdigest()[:16]

    def path_to(self, n):
        out = []
        while n >= 0:
            out.append(self.nodes[n])
            n = self.parents[n]
        out.reverse()
        return out


class RepeatedRRT(RRT):
    """'r-rrt' without pruning: keep the best path, restart the tree at every goal hit."""

    def __init__(self, *args, **kwargs):
        RRT.__init__(self, *args, **kwargs)
        self.bestPath = None
        self.bestPathCost = None

    def reset(self):
        RRT.reset(self)
        self.bestPath = None
        self.bestPathCost = None

    def planMore(self, iters):
        for _ in range(iters):
            self.numIters += 1
            n = self.expand()
            if n is None:
                continue
            if self.goal is not None and self.goal_contains(self.nodes[n]):
                totalcost = self.costs[n] + 0  # objective.terminal == 0
                improved = self.bestPath is None or totalcost < self.bestPathCost
                if improved:
                    self.bestPathCost = totalcost
                    self.bestPath = self.path_to(n)
                RRT.reset(self)  # numIters restarts too, as in the reference (:319,1307)
                return improved
        return False


class RRTForest(object):
    """T independent RRT trials advanced in LOCK STEP: one fused launch per iteration serves all of
    them (`pomp_forest_extend_h`).  Trial t draws from its own `random.Random(seeds[t])`, in the
    reference's order, so its tree is exactly the tree the single-trial planner grows under
    `random.seed(seeds[t])` (pomp/planners/test.py runs such trials one after the other)."""

    def __init__(self, space, start, seeds, goal=None, goal_radius=None, max_step=0.1, resolution=0.01,
                 p_choose_goal=0.1, capacity=16384, device=0, _backends=None):
        self.spec = space_from_pomp(space)
        self.D = self.spec.dim
        self.T = len(seeds)
        self.start = list(start)
        self.goal = None if goal is None else list(goal)
        self.goal_radius = goal_radius
        self.max_step = max_step
        self.resolution = resolution
        self.pChooseGoal = p_choose_goal
        if _backends is None:
            from .backend import CudaForest, CudaSpace
            _backends = (CudaForest, CudaSpace)
        forest_cls, space_cls = _backends
        self._forest = forest_cls(self.D, self.T, capacity, device)
        self._space = space_cls(self.spec, device)
        self.rngs = [random.Random(s) for s in seeds]
        # random numbers drawn ahead for a multi-iteration launch but not used (the trial stopped at the goal):
        # they are the next ones the trial consumes, so every trial sees the reference's sequence
        self._ahead = [collections.deque() for _ in seeds]
        self.multi = hasattr(self._forest, "extend_multi")
        for j in range(self.D):
            self._forest.root[j] = self.start[j]
        self.nodes = [[self.start] for _ in range(self.T)]
        self.parents = [[-1] for _ in range(self.T)]
        self.costs = [[0.0] for _ in range(self.T)]
        self.numIters = [0] * self.T
        self._pending_reset = [True] * self.T
        self._lo_hi = list(zip(self.spec.lo, self.spec.hi))

    def reset_trial(self, t):
        self.nodes[t] = [self.start]
        self.parents[t] = [-1]
        self.costs[t] = [0.0]
        self.numIters[t] = 0
        self._pending_reset[t] = True

    def expand_all(self, active=None):
        """One RRT.expand per active trial; returns the list of new node indices (None where no node
        was added or the trial is inactive)."""
        f, D = self._forest, self.D
        xr = f.xrand
        goal, r, p = self.goal, self.goal_radius, self.pChooseGoal
        for t in range(self.T):
            on = active is None or active[t]
            f.active[t] = 1 if on else 0
            f.reset[t] = 1 if self._pending_reset[t] else 0
            if not on:
                continue
            self._pending_reset[t] = False
            b = t * D
            for j, v in enumerate(self._draw_sample(t)[0]):
                xr[b + j] = v
        f.extend(self._space, self.max_step, self.resolution, True)
        out = [None] * self.T
        added, parent, xnew, elen = f.added, f.parent, f.xnew, f.elen
        for t in range(self.T):
            if f.active[t] and added[t]:
                b = t * D
                par = parent[t]
                self.nodes[t].append([xnew[b + j] for j in range(D)])
                self.parents[t].append(par)
                self.costs[t].append(self.costs[t][par] + elen[t])
                out[t] = len(self.nodes[t]) - 1
        return out

    def _u(self, t):
        a = self._ahead[t]
        return a.popleft() if a else self.rngs[t].random()

    def _draw_sample(self, t):
        """RRT.expand's sample for trial t (kinodynamicplanner.py:358-363): goal-biased with probability
        pChooseGoal, else uniform in the bounds.  random.uniform(a, b) is a + (b - a) * random(); the raw random()
        values are returned as well so that a sample drawn ahead can be handed back."""
        goal, r = self.goal, self.goal_radius
        raw = []
        if goal is not None:
            u0 = self._u(t)
            raw.append(u0)
            if 0.0 + (1.0 - 0.0) * u0 < self.pChooseGoal:
                x = []
                for j in range(self.D):
                    u = self._u(t)
                    raw.append(u)
                    x.append(goal[j] + (-r + (r - (-r)) * u))
                return x, raw
        x = []
        for lo, hi in self._lo_hi:
            u = self._u(t)
            raw.append(u)
            x.append(lo + (hi - lo) * u)
        return x, raw

    def expand_multi(self, M, active=None, stop_at_goal=False):
        """M expansions per active trial in ONE launch (`pomp_forest_extend_multi_h`).  Returns, per trial, the list
        of new node indices (None where an expansion added nothing) of the expansions that were executed: all M,
        or fewer when stop_at_goal and a new node landed in the goal ball.  (Hot host loop: samples are computed
        inline from raw random() values exactly as `_draw_sample` does.)"""
        f, D, T = self._forest, self.D, self.T
        self._ensure_multi_buffers(M)
        xr = f.xrand_m
        goal, r, p, lohi = self.goal, self.goal_radius, self.pChooseGoal, self._lo_hi
        per = D + (1 if goal is not None else 0)
        need = M * per
        raws = [None] * T
        for t in range(T):
            on = active is None or active[t]
            f.active[t] = 1 if on else 0
            f.reset[t] = 1 if (on and self._pending_reset[t]) else 0
            if not on:
                continue
            self._pending_reset[t] = False
            ahead = self._ahead[t]
            rnd = self.rngs[t].random
            if ahead:
                take = min(len(ahead), need)
                raw = [ahead.popleft() for _ in range(take)]
                if take < need:
                    raw += [rnd() for _ in range(need - take)]
            else:
                raw = [rnd() for _ in range(need)]
            raws[t] = raw
            base = t * M * D
            k = 0
            for _ in range(M):
                if goal is not None:
                    u0 = raw[k]
                    k += 1
                    if u0 < p:   # random.uniform(0.0, 1.0) is 0.0 + 1.0 * random(): exactly random()
                        for j in range(D):
                            xr[base + j] = goal[j] + (-r + (r - (-r)) * raw[k + j])
                    else:
                        for j in range(D):
                            lo, hi = lohi[j]
                            xr[base + j] = lo + (hi - lo) * raw[k + j]
                else:
                    for j in range(D):
                        lo, hi = lohi[j]
                        xr[base + j] = lo + (hi - lo) * raw[k + j]
                k += D
                base += D
        f.extend_multi(self._space, M, self.max_step, self.resolution, True,
                       goal if stop_at_goal else None, self.goal_radius if stop_at_goal else 0.0)
        out = [None] * T
        n_done, added_m, parent_m, xnew_m, elen_m = f.n_done, f.added_m, f.parent_m, f.xnew_m, f.elen_m
        for t in range(T):
            raw = raws[t]
            if raw is None:
                continue
            nd = n_done[t]
            if nd < M:
                self._ahead[t].extendleft(reversed(raw[nd * per:]))
            nodes, parents, costs = self.nodes[t], self.parents[t], self.costs[t]
            res = []
            o = t * M
            for _ in range(nd):
                if added_m[o]:
                    par = parent_m[o]
                    b = o * D
                    nodes.append([xnew_m[b + j] for j in range(D)])
                    parents.append(par)
                    costs.append(costs[par] + elen_m[o])
                    res.append(len(nodes) - 1)
                else:
                    res.append(None)
                o += 1
            out[t] = res
        return out

    def _ensure_multi_buffers(self, M):
        f = self._forest
        if hasattr(f, "_alloc_multi"):   # the CPU test double keeps plain lists
            f._alloc_multi()
        elif not hasattr(f, "xrand_m"):
            import ctypes as C
            T, D, MM = self.T, self.D, f.MAX_MULTI
            f.xrand_m = (C.c_double * (T * MM * D))()
            f.n_done = (C.c_int32 * T)()
            f.parent_m = (C.c_int64 * (T * MM))()
            f.xnew_m = (C.c_double * (T * MM * D))()
            f.elen_m = (C.c_double * (T * MM))()
            f.added_m = (C.c_int32 * (T * MM))()
            f._m_alloc = MM

    def goal_contains(self, x):
        s = 0
        for a, b in zip(x, self.goal):
            s = s + (a - b) * (a - b)
        return math.sqrt(s) <= self.goal_radius

    def planMore(self, iters):
        left = iters
        while left > 0:
            M = min(left, 32) if self.multi else 1
            if M > 1:
                for t, res in enumerate(self.expand_multi(M)):
                    self.numIters[t] += len(res)
            else:
                for t in range(self.T):
                    self.numIters[t] += 1
                self.expand_all()
            left -= M

    def trial(self, t):
        """A read-only single-trial view (nodes / parents / getRoadmap / roadmap_hash) of trial t."""
        v = RRT.__new__(RRT)
        v.nodes, v.parents, v.costs = self.nodes[t], self.parents[t], self.costs[t]
        return v


class RepeatedRRTForest(RRTForest):
    """'r-rrt' trials in lock step: per trial, keep the best path cost and restart the tree at every
    goal hit (RepeatedRRT.planMore, kinodynamicplanner.py:1291-1312)."""

    def __init__(self, *args, **kwargs):
        RRTForest.__init__(self, *args, **kwargs)
        self.bestPathCost = [None] * self.T
        self.bestPath = [None] * self.T
        self.totalIters = [0] * self.T

    def planMore(self, iters):
        """Per trial exactly RepeatedRRT.planMore(iters): a trial that reaches the goal RETURNS from this call
        (kinodynamicplanner.py:1291-1312) -- here it sits out the rest of the call -- so under the reference
        harness's planMore(10) loop (planners/test.py:24-40) trial t reproduces the single-trial run of its seed
        iteration for iteration."""
        live = [True] * self.T
        left = iters
        while left > 0 and any(live):
            M = min(left, 32) if self.multi else 1
            if M > 1:
                new_lists = self.expand_multi(M, live, stop_at_goal=self.goal is not None)
            else:
                new_lists = [None if n is None and not live[t] else [n] for t, n in enumerate(self.expand_all(live))]
            left -= M
            for t, res in enumerate(new_lists):
                if not live[t] or res is None:
                    continue
                self.numIters[t] += len(res)
                self.totalIters[t] += len(res)
                n = res[-1] if res else None
                if n is None:
                    continue
                # only the LAST executed expansion can have reached the goal (the launch stops there)
                if self.goal is not None and self.goal_contains(self.nodes[t][n]):
                    totalcost = self.costs[t][n] + 0
                    if self.bestPath[t] is None or totalcost < self.bestPathCost[t]:
                        self.bestPathCost[t] = totalcost
                        path, i = [], n
                        while i >= 0:
                            path.append(self.nodes[t][i])
                            i = self.parents[t][i]
                        path.reverse()
                        self.bestPath[t] = path
                    self.reset_trial(t)
                    live[t] = False


def testForest(forest, max_iters, filename, check_every=10, trial_offset=0):
    """The lock-step counterpart of pomp.planners.test.testPlanner (pomp/planners/test.py:8-52): all trials of
    `forest` (a RepeatedRRTForest) advance together for `max_iters` planner iterations, and the result file has
    the reference's layout byte for byte -- header ``trial,plan iters,plan time,best cost``, one row per change of
    a trial's best cost (sampled every `check_every` = 10 iterations, like the reference's planMore(10) loop), one
    closing row per trial -- so processresults.py / viewsummary.py read GPU runs unchanged.  Rows are grouped by
    trial, trials numbered from `trial_offset` (ranks of a multi-GPU run write disjoint trial ranges).  The time
    column is the wall time of the lock-step run at that point; the closing row carries the total.  Returns the
    rows as (trial, iters, time, cost) tuples."""
    import time
    T = forest.T
    rows = [[] for _ in range(T)]
    cur = [float("inf")] * T
    t0 = time.time()
    iters = 0
    while iters < max_iters:
        step = min(check_every, max_iters - iters)
        forest.planMore(step)
        iters += step
        t1 = time.time()
        for t in range(T):
            c = forest.bestPathCost[t]
            if c is not None and c != cur[t]:
                cur[t] = c
                rows[t].append((trial_offset + t, iters, t1 - t0, c))
    total = time.time() - t0
    out = []
    with open(filename, "w") as f:
        f.write("trial,plan iters,plan time,best cost\n")
        for t in range(T):
            for r in rows[t] + [(trial_offset + t, iters, total, cur[t])]:
                f.write(str(r[0]) + "," + str(r[1]) + "," + str(r[2]) + "," + str(r[3]) + "\n")
                out.append(r)
    return out
