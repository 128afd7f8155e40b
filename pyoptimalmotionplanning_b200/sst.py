"""Lock-step Stable-Sparse-RRT* forest: T independent, seeded sst* trials advanced together on the GPU
(BASELINE config 5: `DoubleIntegrator sst*`, 4096 seeded trials spread over the GPUs).

What the reference does per trial and iteration (pomp/planners/rrtstarplanner.py):

  StableSparseRRTStar.planMore  :454-471   restart schedule, radii *= shrinkage
  StableSparseRRT.expand        :300-343   sample, pickNode, control, edge check, cost, add, prune
  StableSparseRRT.pickNode      :390-410   neighbors(selectionRadius) -> cheapest ACTIVE node, else nearest active
  StableSparseRRT.nodeLocallyBest :345-366 nearest witness / witnesses within witnessRadius
  StableSparseRRT.doPruning     :368-388   nearest witness gets the new representative, leaf chain removed
  CVControlSpace.trajectory     spaces/statespace.py:47-61, EpsilonEdgeChecker spaces/edgechecker.py:20-30

Division of labour.  Everything that touches the tree -- the radius / nearest scans over the trial's nodes and
witnesses, the stepwise trajectory, the sampled edge check, costs, witness bookkeeping, pruning, the goal test --
runs on the device, one block per trial, several iterations per launch (csrc/sst.cu).  Everything that touches
Python's `random` stays on the host and is drawn from THE CALLER'S OWN OBJECTS in the reference's order:
`random.uniform(0,1)` for the goal bias, `goalSampler.sample()` or `configurationSampler.sample()`, then
`controlSpace.controlSet(x).sample()` / `.contains(u)` (RandomControlSelector with numSamples = 1,
kinodynamicplanner.py:66-84).  Each trial owns a `random.Random(seed)`; while a trial's samples are drawn the
functions of the global `random` module are pointed at that generator, so trial t consumes exactly the stream the
reference consumes after `random.seed(seed_t)`.  Two assumptions make the pre-drawing legal, both hold for the
KinodynamicSpace family: the control set does not depend on the state, and pickNode never returns None (there is
always an active node); the device counts violations of the second (`n_none`) and the driver raises on any.

The restart schedule of sst* (:462-470) depends only on the iteration count, so it is computed here exactly as the
reference does and handed to the device as per-iteration radii + restart flags.
"""
import math
import random as _random

import numpy as np

from . import _abi
from .descriptors import DynamicsSpec, MetricSpec, dynamics_from_pomp, geodesic_of_space, space_from_pomp

_RANDOM_FUNCS = ("random", "uniform", "gauss", "randint", "choice", "randrange", "sample", "shuffle", "normalvariate",
                 "triangular", "betavariate", "expovariate", "gammavariate", "lognormvariate", "vonmisesvariate",
                 "paretovariate", "weibullvariate", "getrandbits")


class _StreamSwitch(object):
    """Points the functions of the global `random` module at one trial's generator at a time."""

    def __init__(self):
        self._saved = None

    def __enter__(self):
        self._saved = {n: getattr(_random, n) for n in _RANDOM_FUNCS if hasattr(_random, n)}
        return self

    def use(self, rng):
        for n in self._saved:
            setattr(_random, n, getattr(rng, n))

    def __exit__(self, *exc):
        for n, f in self._saved.items():
            setattr(_random, n, f)
        self._saved = None


def _default_backend(x_dim, u_dim, n_trials, node_cap, witness_cap, dyn, geo, space_spec, device):
    from .backend import CudaSSTForest
    return CudaSSTForest(x_dim, u_dim, n_trials, node_cap, witness_cap, dyn, geo, space_spec, device)


class SSTStarForest(object):
    """T seeded sst* trials in lock step.

    control_space: the caller's pomp ControlSpace (KinodynamicSpace family, e.g. CVControlSpace)
    start, goal:   start state; goal = the caller's goal Set (NeighborhoodSubset) or None
    seeds:         one per trial; trial t reproduces the reference run after random.seed(seeds[t])
    objective:     TimeObjectiveFunction (incremental cost = u[0], terminal 0)
    """

    MAX_MULTI = 32

    def __init__(self, control_space, start, goal, seeds, objective=None, selectionRadius=0.2, witnessRadius=0.2,
                 numSSTIters=300, shrinkage=0.8, pChooseGoal=0.1, resolution=0.01, node_capacity=4096,
                 witness_capacity=4096, device=0, configurationSampler=None, _backend_factory=None):
        self.controlSpace = control_space
        self.cspace = control_space.configurationSpace()
        self.start = list(start)
        self.goal = goal
        self.seeds = list(seeds)
        self.T = len(self.seeds)
        if objective is not None and type(objective).__name__ != "TimeObjectiveFunction":
            raise TypeError("SSTStarForest supports TimeObjectiveFunction (cost = duration), not %s"
                            % type(objective).__name__)
        self.selectionRadius0, self.witnessRadius0 = float(selectionRadius), float(witnessRadius)
        self.numIters0, self.shrinkage = numSSTIters, shrinkage
        self.pChooseGoal = pChooseGoal
        self.resolution = resolution
        self.dyn = dynamics_from_pomp(control_space, stepwise=True)
        if self.dyn.cost_kind != _abi.COST_NONE:
            self.dyn = self.dyn.with_cost(_abi.COST_NONE)
        X = self.dyn.x_dim
        self.X, self.U = X, self.dyn.u_dim
        self.geo = geodesic_of_space(self.cspace, X)
        self.space_spec = space_from_pomp(self.cspace)
        if goal is not None:
            if type(goal).__name__ != "NeighborhoodSubset":
                raise TypeError("goal must be a NeighborhoodSubset (centre + radius)")
            self.goal_c, self.goal_r = [float(v) for v in goal.c], float(goal.r)
        else:
            self.goal_c, self.goal_r = [0.0] * X, -1.0
        self._rngs = [_random.Random(s) for s in self.seeds]
        if configurationSampler is None:
            cs = self.cspace
            configurationSampler = type("Sampler", (), {"sample": lambda self_: cs.sample()})()
        self.configurationSampler = configurationSampler
        self._goal_sample = None if goal is None else self._make_goal_sampler()
        self._be = (_backend_factory or _default_backend)(X, self.U, self.T, node_capacity, witness_capacity, self.dyn,
                                                          self.geo, self.space_spec, device)
        self.reset()

    def _make_goal_sampler(self):
        # SubsetSampler.sample (spaces/sampler.py:33-41)
        space, subset = self.cspace, self.goal

        def sample():
            res = subset.sample()
            if res is not None:
                return res
            for _ in range(1000):
                x = space.sample()
                if subset.contains(x):
                    return x
            return space.sample()
        return sample

    # ------------------------------------------------------------------ StableSparseRRTStar.reset
    def reset(self):
        self.numIters = 0
        self.restartCount = 0
        self.itersleft = self.numIters0
        self.selectionRadius, self.witnessRadius = self.selectionRadius0, self.witnessRadius0
        self.bestPathCost = [float("inf")] * self.T
        self.numNodes = [1] * self.T
        self._pending_restart = False
        self._be.reset(self.start)

    # ------------------------------------------------------------------ planMore for every trial
    def planMore(self, iters):
        """Advances every trial by `iters` iterations.  Returns the list of trials whose best cost improved."""
        improved = set()
        done = 0
        while done < iters:
            M = min(self.MAX_MULTI, iters - done)
            xr, uu, valid = self._draw(M)
            sel, wit, restart = [], [], []
            for _ in range(M):
                # the iteration runs with the current radii; the restart bookkeeping follows it (:454-471)
                sel.append(self.selectionRadius)
                wit.append(self.witnessRadius)
                restart.append(1 if self._pending_restart else 0)
                self._pending_restart = False
                self.numIters += 1
                self.itersleft -= 1
                if self.itersleft <= 0:
                    self.restartCount += 1
                    self._pending_restart = True          # ssrrt.reset() before the next iteration
                    self.selectionRadius *= self.shrinkage
                    self.witnessRadius *= self.shrinkage
                    self.itersleft = ((1.0 + math.log(self.restartCount))
                                      * pow(self.shrinkage, -(len(self.start) + 1) * self.restartCount) * self.numIters0)
            best, nodes, none = self._be.step(M, xr, uu, valid, sel, wit, restart, self.goal_c, self.goal_r,
                                              self.resolution, half_sq=self._half_squares(uu))
            if any(none):
                raise RuntimeError("pickNode found no active node in trial(s) %s: the pre-drawn random stream is out "
                                   "of step with the reference" % [t for t, c in enumerate(none) if c])
            for t in range(self.T):
                if best[t] < self.bestPathCost[t]:
                    self.bestPathCost[t] = float(best[t])
                    improved.add(t)
                self.numNodes[t] = int(nodes[t])
            done += M
        if self._pending_restart:
            # the reference resets the tree right after the iteration that exhausted the budget: make it visible
            self._be.step(0, np.zeros((self.T, 0, self.X)), np.zeros((self.T, 0, self.U)), np.zeros((self.T, 0), np.uint8),
                          [], [], [], self.goal_c, self.goal_r, self.resolution, restart_now=True)
            self._pending_restart = False
            self.numNodes = [1] * self.T
        return sorted(improved)

    def _draw(self, M):
        """The next M iterations' samples of every trial, from the caller's samplers in the reference's order."""
        T, X, U = self.T, self.X, self.U
        xr = np.empty((T, M, X))
        uu = np.empty((T, M, U))
        valid = np.ones((T, M), np.uint8)
        x0 = self.start
        with _StreamSwitch() as sw:
            for t in range(T):
                sw.use(self._rngs[t])
                for m in range(M):
                    if self._goal_sample is not None and _random.uniform(0.0, 1.0) < self.pChooseGoal:
                        xr[t, m] = self._goal_sample()
                    else:
                        xr[t, m] = self.configurationSampler.sample()
                    Uset = self.controlSpace.controlSet(x0)   # (TimeBiasSet's constructor draws, like the reference)
                    u = Uset.sample()
                    uu[t, m] = u
                    if not Uset.contains(u):
                        valid[t, m] = 0
        return xr, uu, valid

    def _half_squares(self, uu):
        """CVControlSpace.trajectory (statespace.py:55-59) multiplies by 0.5*dt**2; Python's ** is libm pow, which is
        not always dt*dt, and a node coordinate that is off by one ulp is a different tree.  The step lengths depend
        on the duration only, so the two coefficients (full step, last partial step) are computed here, with Python's
        own arithmetic, for every pre-drawn control."""
        if self.dyn.kind != _abi.DYN_CV_EULER:
            return None
        dt0 = self.dyn.dt
        T, M = uu.shape[0], uu.shape[1]
        last = np.zeros((T, M))
        for t in range(T):
            for m in range(M):
                duration = float(uu[t, m, 0])
                tt, dl = 0.0, dt0
                while tt < duration:
                    dl = min(dt0, duration - tt)
                    tt = min(tt + dt0, duration)
                last[t, m] = 0.5 * dl ** 2
        return last, 0.5 * dt0 ** 2

    # ------------------------------------------------------------------ results
    def getRoadmap(self, trial):
        """(V, E) of TreePlanner.getRoadmap (kinodynamicplanner.py:211-238) for one trial: the same traversal over
        the same children order, E without the controls: [(i, j)]."""
        x, parent, alive = self._be.get_trial(trial)
        n = len(x)
        children = [[] for _ in range(n)]
        for i in range(1, n):
            if alive[i] and parent[i] >= 0:
                children[parent[i]].append(i)         # creation order = order of addChild
        V, E = [], []
        if n == 0 or not alive[0]:
            return V, E
        V.append(list(self.start))            # the reference keeps the caller's own start list (ints stay ints)
        q = [(0, 0)]
        while q:
            node, i = q.pop()
            for c in children[node]:
                j = len(V)
                E.append((i, j))
                V.append(list(x[c]))
                q.append((c, j))
        return V, E
