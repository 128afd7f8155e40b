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
