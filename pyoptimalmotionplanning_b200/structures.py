"""Drop-in for pomp.structures.nearestneighbors.NearestNeighbors (nearestneighbors.py:13-141).

Same constructor and the same seven methods; points and payloads are kept BY REFERENCE on the
host (planners compare `n.x` by value in remove() and expect nearest() to hand back the very
objects they stored), coordinates live on the GPU in insertion order.

Semantics chosen where the reference's two back ends disagree:
  * the exact nearest neighbour under the metric is returned (= method 'bruteforce'; the kd-tree
    is only approximate on SO(2)/cost metrics);
  * ties: lowest insertion index wins; knearest is ordered by (distance, insertion index);
    neighbors is ordered by insertion index;
  * neighbors uses d <= radius, like the reference's default kd-tree back end (kdtree.py:347);
    set `inclusive_radius = False` for the brute-force back end's d < radius (:139);
  * nearest() on an empty set returns None (brute-force behaviour, :88-96).
"""
from __future__ import print_function

import numpy as np

from .descriptors import MetricSpec, metric_from_pomp

_COMPACT_MIN = 1024


def _default_backend(spec, device):
    from .backend import CudaNN
    return CudaNN(spec, device)


class NearestNeighbors(object):
    def __init__(self, metric, method="b200", device=0, _backend_factory=None):
        self.metric = metric
        self.method = method
        self.device = device
        self.inclusive_radius = True
        self._factory = _backend_factory or _default_backend
        self._be = None
        self._sig = None
        self._dim = metric.dim if isinstance(metric, MetricSpec) else None
        self._rewire_checker = None
        self.generation = 0   # bumped whenever insertion indices are renumbered (set(), reset(), compaction)
        self._clear_host()

    def enable_rewire_prefetch(self, checker):
        """RRT* rewire batch (rrtstarplanner.py:122-187): `checker` is this package's EpsilonEdgeChecker.
        From now on knearest(pt, k) also decides, in the same device call, the straight edges between pt and
        each of its k neighbours (both directions); the checker serves the planner's following
        feasible(LinearInterpolator) calls from that batch.  Results are identical (edge feasibility is a
        pure function of the endpoints); pass None to switch it off."""
        self._rewire_checker = checker

    # ------------------------------------------------------------------ host bookkeeping
    def _clear_host(self):
        self._points = []   # insertion index -> point object
        self._datas = []    # insertion index -> payload
        self._alive = []
        self._ndead = 0
        self._lookup = {}   # tuple(point) -> [indices], insertion order

    def _metric_signature(self):
        """Cheap fingerprint of the mutable parts of the metric (CostSpaceRRT.updateBestCost mutates
        costMax / costWeight in place: kinodynamicplanner.py:1059-1077)."""
        m = self.metric
        if isinstance(m, MetricSpec):
            return m.key()
        sig = [getattr(m, "costWeight", None), getattr(m, "costMax", None)]
        base = getattr(m, "metric", m)
        owner = getattr(base, "__self__", None)
        geo = getattr(owner, "geodesic", None)
        w = getattr(geo, "componentWeights", None)
        if w is not None:
            sig.append(tuple(w))
        w = getattr(base, "w", None)
        if w is not None:
            sig.append(tuple(w) if hasattr(w, "__iter__") else w)
        return tuple(sig)

    def _ensure_backend(self, dim):
        if self._be is None:
            self._dim = dim
            spec = metric_from_pomp(self.metric, dim)
            if spec.dim != dim:
                raise ValueError("metric is %d-dimensional but points have %d coordinates" % (spec.dim, dim))
            self._be = self._factory(spec, self.device)
            self._sig = self._metric_signature()
        return self._be

    def _sync_metric(self):
        sig = self._metric_signature()
        if sig != self._sig:
            self._be.update_metric(metric_from_pomp(self.metric, self._dim))
            self._sig = sig

    def __len__(self):
        return len(self._points) - self._ndead

    # ------------------------------------------------------------------ reference API
    def reset(self):
        self.generation += 1
        self._clear_host()
        if self._be is not None:
            self._be.reset()

    def add(self, point, data=None):
        """Adds a point with an associated datum."""
        be = self._ensure_backend(len(point))
        idx = be.add_one(point)
        assert idx == len(self._points)
        self._points.append(point)
        self._datas.append(data)
        self._alive.append(True)
        self._lookup.setdefault(tuple(point), []).append(idx)

    def remove(self, point, data=None):
        """Removes the first stored entry with p == point (and pd == data when data is given).
        Returns the number of points removed (0 or 1)."""
        key = tuple(point)
        cands = self._lookup.get(key)
        if cands:
            for pos, idx in enumerate(cands):
                if data is None or self._datas[idx] == data:
                    del cands[pos]
                    if not cands:
                        del self._lookup[key]
                    self._alive[idx] = False
                    self._ndead += 1
                    self._be.remove(idx)
                    self._maybe_compact()
                    return 1
        print("NearestNeighbors: unable to remove", point, "does not exist")
        return 0

    def set(self, points, datas=None):
        """Sets the point set to a list of points."""
        points = list(points)
        if datas is None:
            datas = [None] * len(points)
        datas = list(datas)
        self.generation += 1
        self._clear_host()
        if self._be is not None:
            self._be.reset()
        if not points:
            return
        be = self._ensure_backend(len(points[0]))
        be.set(np.asarray(points, dtype=np.float64))
        self._points = points
        self._datas = datas
        self._alive = [True] * len(points)
        for i, p in enumerate(points):
            self._lookup.setdefault(tuple(p), []).append(i)

    def nearest(self, pt, filter=None):
        """Nearest neighbor lookup, possibly with filter (filter(point,data) == True rejects)."""
        if self._be is None or len(self) == 0:
            return None
        self._sync_metric()
        idx, _ = self._be.nearest_one(pt)
        if idx < 0:
            return None
        if filter is None or not filter(self._points[idx], self._datas[idx]):
            return (self._points[idx], self._datas[idx])
        # the winner was rejected: walk the candidates in (distance, index) order
        n = len(self._points)
        k = 8
        while True:
            k = min(k, 128, n)
            ids, _, cnt = self._be.knearest(np.asarray([pt], dtype=np.float64), k)
            for j in range(int(cnt[0])):
                i = int(ids[0, j])
                if not filter(self._points[i], self._datas[i]):
                    return (self._points[i], self._datas[i])
            if int(cnt[0]) < k or k >= n:
                return None
            if k >= 128:
                break
            k *= 4
        # more than 128 rejected in a row: evaluate the predicate everywhere and mask on the device
        return self._masked(filter, lambda: self._first(self._be.nearest_one(pt)))

    def knearest(self, pt, k, filter=None):
        """K-nearest neighbor lookup, possibly with filter.  Sorted by (distance, insertion index)."""
        if self._be is None or len(self) == 0:
            return []
        self._sync_metric()
        base = None
        if filter is not None:
            base = np.zeros(len(self._points), dtype=np.uint8)
            for i, (p, d) in enumerate(zip(self._points, self._datas)):
                if self._alive[i] and filter(p, d):
                    base[i] = 1
        q = np.asarray([pt], dtype=np.float64)
        out = []
        try:
            if base is not None:
                self._be.set_mask(base)
            if k <= 128:
                chk = self._rewire_checker
                if chk is not None and base is None and chk.prefetch_ready(self._be):
                    # RRT* rewire batch: the same k nearest plus both straight edges per neighbour in ONE call;
                    # the checker answers the planner's following feasible() calls from this batch
                    ids, _, ok_in, ok_out = self._be.rewire_candidates(chk._dev._be, pt, k, chk.resolution)
                    res = [(self._points[int(i)], self._datas[int(i)]) for i in ids]
                    chk.prime(pt, [r[0] for r in res], ok_in, ok_out)
                    return res
                ids, _, cnt = self._be.knearest(q, k)
                return [(self._points[int(i)], self._datas[int(i)]) for i in ids[0, :int(cnt[0])]]
            # very large k: peel off 128 at a time by masking what was already returned
            mask = base if base is not None else np.zeros(len(self._points), dtype=np.uint8)
            while len(out) < k:
                ids, _, cnt = self._be.knearest(q, min(128, k - len(out)))
                c = int(cnt[0])
                if c == 0:
                    break
                for i in ids[0, :c]:
                    out.append((self._points[int(i)], self._datas[int(i)]))
                    mask[int(i)] = 1
                self._be.set_mask(mask)
                base = mask
        finally:
            if base is not None:
                self._be.set_mask(None)
        return out

    def neighbors(self, pt, radius):
        """Range query, all points within radius of pt, in insertion order."""
        if self._be is None or len(self) == 0:
            return []
        self._sync_metric()
        off, ids = self._be.neighbors(np.asarray([pt], dtype=np.float64), radius, self.inclusive_radius)
        return [(self._points[int(i)], self._datas[int(i)]) for i in ids]

    def density(self, pt, radius, sigma):
        """EST.density (kinodynamicplanner.py:688-693): sum of exp(-(d/sigma)^2) over the nodes within radius."""
        if self._be is None or len(self) == 0:
            return 0.0
        self._sync_metric()
        return float(self._be.density(np.asarray([pt], dtype=np.float64), radius, sigma, self.inclusive_radius)[0][0])

    def density_batch(self, queries, radius, sigma):
        self._sync_metric()
        return self._be.density(queries, radius, sigma, self.inclusive_radius)

    # ------------------------------------------------------------------ batched extras
    def nearest_batch(self, queries):
        """(idx[nq], dist[nq]) for a batch of queries; idx = -1 where nothing is admissible."""
        self._sync_metric()
        return self._be.nearest(queries)

    def knearest_batch(self, queries, k):
        self._sync_metric()
        return self._be.knearest(queries, k)

    def neighbors_batch(self, queries, radius):
        """CSR (offsets[nq+1], idx[hits])."""
        self._sync_metric()
        return self._be.neighbors(queries, radius, self.inclusive_radius)

    def item(self, index):
        return (self._points[index], self._datas[index])

    # ------------------------------------------------------------------ helpers
    def _first(self, res):
        idx = res[0]
        return None if idx < 0 else (self._points[idx], self._datas[idx])

    def _masked(self, filter, fn):
        mask = np.zeros(len(self._points), dtype=np.uint8)
        for i, (p, d) in enumerate(zip(self._points, self._datas)):
            if self._alive[i] and filter(p, d):
                mask[i] = 1
        self._be.set_mask(mask)
        try:
            return fn()
        finally:
            self._be.set_mask(None)

    def _maybe_compact(self):
        """Once half of the entries are dead the set is rebuilt from the live ones.  Insertion indices handed out
        by the *_batch methods / item() are renumbered by that: they are valid only while `generation` is
        unchanged (compare it before and after a remove()); set `auto_compact = False` to keep indices stable."""
        if not getattr(self, "auto_compact", True):
            return
        n = len(self._points)
        if n < _COMPACT_MIN or self._ndead * 2 < n:
            return
        keep = [i for i in range(n) if self._alive[i]]
        pts = [self._points[i] for i in keep]
        datas = [self._datas[i] for i in keep]
        self.set(pts, datas)


def _density_backend(dim, bases, device):
    from .backend import CudaProjectionDensity
    return CudaProjectionDensity(dim, bases, device)


class ProjectionHashDensity(object):
    """The density estimator of ESTWithProjections (kinodynamicplanner.py:746-868) with the node set on the GPU.

    Mirrors the reference's bookkeeping: generateDefaultBases(indices) (:765-815) enumerates all subsets of
    min(len(indices), 4) coordinates, one single-coordinate basis vector {i: resolution*scale_i} per member, and repeats
    them `hierarchyLevels` times with the scales multiplied by (1 + level); onAddNode(x) (:817-832) files the node
    under its integer cell per basis; density(x) (:834-856) averages the (level-weighted) cell populations.  Here the
    nodes are coordinates on the device and the populations are counted by a kernel (csrc/est.cu); values are
    bit-identical to the reference's.

    bounds: (bmin, bmax) of the configuration space (controlSpace.configurationSpace().bounds()) or None (scale 1);
    cost_space: True when the planner runs on a CostControlSpace (the last scale is doubled, :772-773).
    """

    def __init__(self, dim, bounds=None, resolution=13, hierarchyLevels=2, cost_space=False, indices=None, device=0,
                 _backend_factory=None):
        import itertools
        self.dim = dim
        self.projectionResolution = resolution
        self.projectionHierarchyLevels = hierarchyLevels
        if bounds is not None:
            scale = [1.0 / (b - a) for (a, b) in zip(bounds[0], bounds[1])]
            if cost_space:
                scale[-1] *= 2.0
        else:
            scale = [1] * dim
        indices = list(range(dim)) if indices is None else list(indices)
        d = min(len(indices), 4)
        base = []
        for element in itertools.combinations(indices, d):
            base.append([(i, resolution * scale[i]) for i in element])
        bases = list(base)
        for level in range(1, hierarchyLevels):
            for b in base:
                bases.append([(i, v * (1 + level)) for i, v in b])
        n = len(bases)
        baselevels = n / hierarchyLevels
        # density(): blevel = int(index / baselevels) + 1, contribution cnt * pow(blevel, len(basisvector)) with
        # single-entry basis vectors
        self.bases = [([i for i, _ in b], [v for _, v in b], float(pow(int(j / baselevels) + 1, 1)))
                      for j, b in enumerate(bases)]
        self._be = (_backend_factory or _density_backend)(dim, self.bases, device)

    def reset(self):
        self._be.reset()

    def __len__(self):
        return int(self._be.size())

    def onAddNode(self, x):
        self._be.add([list(x)])

    def add_batch(self, xs):
        self._be.add(xs)

    def density(self, x):
        return float(self._be.density([list(x)])[0])

    def density_batch(self, xs):
        return self._be.density(xs)
