"""Drop-ins for the feasibility / edge-check / propagation objects of pomp.spaces.

  DeviceSpace            ConfigurationSpace.feasible          (configurationspace.py:46-48,142,196-199)
  EpsilonEdgeChecker     EpsilonEdgeChecker.feasible          (edgechecker.py:11-30) + visible(a,b)
  DeviceControlSpace     ControlSpace.nextState               (controlspace.py:25-27 and subclasses)
  BatchedRandomControlSelector  RandomControlSelector.select  (kinodynamicplanner.py:57-84)

Every object accepts either the unmodified pomp object (translated by duck typing, see
descriptors.py) or the explicit *Spec.  Anything that cannot be translated raises TypeError:
there is no CPU path to fall back to.
"""
import math

import numpy as np

from . import _abi
from .descriptors import (DynamicsSpec, MetricSpec, SpaceSpec, dynamics_from_pomp, geodesic_of_space,
                          metric_from_pomp, space_from_pomp, _geodesic_components)


def _space_backend(spec, device):
    from .backend import CudaSpace
    return CudaSpace(spec, device)


def _dyn_backend(spec):
    from .backend import CudaDynamics
    return CudaDynamics(spec)


class DeviceSpace(object):
    """GPU-side feasible() of a pomp ConfigurationSpace (or a SpaceSpec)."""

    def __init__(self, space, device=0, _backend_factory=None):
        self.space = space
        self.spec = space_from_pomp(space)
        self._factory = _backend_factory or _space_backend
        self._device = device
        self._be = self._factory(self.spec, device)
        self._tracks_pomp = not isinstance(space, SpaceSpec)
        self._obstacle_fp = self._bounds_and_fingerprint(space)[2] if self._tracks_pomp else None

    def dimension(self):
        return self.spec.dim

    @staticmethod
    def _bounds_and_fingerprint(space):
        """(lo, hi, fingerprint of the obstacle lists) read WITHOUT re-translating the obstacles: O(components),
        not O(obstacles), per call."""
        comps = getattr(space, "components", None)
        lo, hi, fp = [], [], []
        for c in (comps if comps is not None else [space]):
            box = getattr(c, "box", None)
            if box is not None and hasattr(box, "bmin"):
                lo += list(box.bmin)
                hi += list(box.bmax)
            else:
                d = c.dimension()
                lo += [-math.inf] * d
                hi += [math.inf] * d
            obs = getattr(c, "obstacles", None)
            if obs is not None:
                fp.append((id(obs), len(obs)))
        return lo, hi, tuple(fp)

    def _refresh_bounds(self):
        # CostControlSpace.setCostMax mutates bounds in place (costspace.py:26-30); an addObstacle() after
        # construction (geometric.py:62-63) changes the obstacle list: the device tables are rebuilt then
        if not self._tracks_pomp:
            return
        lo, hi, fp = self._bounds_and_fingerprint(self.space)
        if fp != self._obstacle_fp:
            self.spec = space_from_pomp(self.space)
            self._be = self._factory(self.spec, self._device)
            self._obstacle_fp = fp
            self.bounds_version = getattr(self, "bounds_version", 0) + 1
            return
        lo = [float(v) for v in lo]
        hi = [float(v) for v in hi]
        if lo != self.spec.lo or hi != self.spec.hi:
            self.spec.lo, self.spec.hi = lo, hi
            self._be.set_bounds(lo, hi)
            self.bounds_version = getattr(self, "bounds_version", 0) + 1

    def feasible(self, x):
        self._refresh_bounds()
        return self._be.feasible_one(x)

    def contains(self, x):
        return self.feasible(x)

    def feasible_batch(self, xs):
        self._refresh_bounds()
        return self._be.feasible(xs)


class EpsilonEdgeChecker(object):
    """Fixed-resolution sampled edge checker on the GPU; same constructor and feasible() as
    pomp.spaces.edgechecker.EpsilonEdgeChecker."""

    def __init__(self, space, resolution, device=0, _backend_factory=None):
        self.space = space
        self.resolution = resolution
        self._dev = space if isinstance(space, DeviceSpace) else DeviceSpace(space, device, _backend_factory)
        self._geo_cache = {}
        self._rewire = None      # (resolution, {(tuple(a), tuple(b)): feasible}) of the last rewire batch
        self.rewire_hits = 0

    # ---- RRT* rewire batch (filled by NearestNeighbors.knearest, see enable_rewire_prefetch)
    def prefetch_ready(self, nn_backend):
        """The batch call needs both back ends on the device library (or both test fakes)."""
        self._dev._refresh_bounds()
        be = self._dev._be
        if not hasattr(nn_backend, "rewire_candidates") or be is None:
            return False
        # pomp_rewire_candidates_h needs the space and the node set in the same coordinates on the same device
        # (a CostMetric / state-lifted NearestNeighbors has dim + 1): otherwise the plain k-NN path answers
        if getattr(nn_backend, "dim", None) != self._dev.spec.dim:
            return False
        return getattr(nn_backend, "device", 0) == getattr(be, "device", 0)

    def prime(self, x, neighbours, ok_in, ok_out):
        tx = tuple(x)
        cache = {}
        for p, a, b in zip(neighbours, ok_in, ok_out):
            tp = tuple(p)
            cache[(tp, tx)] = bool(a)
            cache[(tx, tp)] = bool(b)
        self._rewire = (self.resolution, cache, getattr(self._dev, "bounds_version", 0))

    # ---- reference API
    def feasible(self, interpolator):
        self._dev._refresh_bounds()
        be = self._dev._be
        name = type(interpolator).__name__
        if name == "LinearInterpolator":  # interpolators.py:28-35
            if (self._rewire is not None and self._rewire[0] == self.resolution
                    and self._rewire[2] == getattr(self._dev, "bounds_version", 0)):
                hit = self._rewire[1].get((tuple(interpolator.a), tuple(interpolator.b)))
                if hit is not None:
                    self.rewire_hits += 1
                    return hit
            return be.edge_check_one(interpolator.a, interpolator.b, self.resolution)
        if name == "PiecewiseLinearInterpolator" and interpolator.trajectory is None:  # :73-102
            path = interpolator.path
            g = self._geodesic_spec(interpolator.geodesic)
            return bool(be.edge_check_path(g, np.asarray(path, dtype=np.float64), [0, len(path)],
                                           self.resolution)[0])
        if name == "GeodesicInterpolator":  # :37-48
            g = self._geodesic_spec(interpolator.space)
            return bool(be.edge_check_path(g, np.asarray([interpolator.a, interpolator.b], dtype=np.float64),
                                           [0, 2], self.resolution)[0])
        if name == "DubinsCarInterpolator":  # example_problems/dubins.py:62-73
            dyn = DynamicsSpec(_abi.DYN_DUBINS, 3, 2)
            return bool(be.edge_check_control(dyn, None, [interpolator.x], [interpolator.control],
                                              self.resolution)[0])
        # unknown curve (LambdaInterpolator, ...): the caller's eval() runs on the host, the
        # feasibility of every sample of the reference's schedule is decided on the GPU
        return self._generic(interpolator)

    def visible(self, a, b):
        """Klamp't-style visibility: the straight edge a->b (space.interpolator(a,b)) is feasible."""
        self._dev._refresh_bounds()
        return self._dev._be.edge_check_one(a, b, self.resolution)

    # ---- batched extras
    def feasible_batch(self, a, b):
        """Linear edges a[i]->b[i]."""
        self._dev._refresh_bounds()
        return self._dev._be.edge_check(a, b, self.resolution)

    def feasible_paths(self, waypoints, offsets, geodesic=None):
        g = geodesic if isinstance(geodesic, MetricSpec) else self._geodesic_spec(geodesic)
        return self._dev._be.edge_check_path(g, waypoints, offsets, self.resolution)

    def feasible_controls(self, control_space, x, u):
        """Edges controlSpace.interpolator(x[i],u[i]) (base states, no cost coordinate)."""
        dyn = dynamics_from_pomp(control_space, stepwise=True)
        if dyn.cost_kind != _abi.COST_NONE:
            dyn = dyn.with_cost(_abi.COST_NONE)
        g = None
        if dyn.kind in (_abi.DYN_PENDULUM, _abi.DYN_CV_EULER):
            cs = control_space.configurationSpace() if hasattr(control_space, "configurationSpace") else None
            g = geodesic_of_space(cs, dyn.x_dim) if cs is not None else MetricSpec.euclidean(dyn.x_dim)
        return self._dev._be.edge_check_control(dyn, g, x, u, self.resolution)

    # ---- helpers
    def _geodesic_spec(self, geodesic):
        if geodesic is None:
            return MetricSpec.euclidean(self._dev.spec.dim)
        key = id(geodesic)
        w = tuple(getattr(geodesic, "componentWeights", ()))
        hit = self._geo_cache.get(key)
        if hit is not None and hit[0] == w:
            return hit[1]
        geo = getattr(geodesic, "geodesic", geodesic)
        if hasattr(geo, "components") and hasattr(geo, "componentWeights"):
            comps, ww = _geodesic_components(geo)
            spec = MetricSpec.geodesic_product(comps, ww)
        else:
            mro = [k.__name__ for k in type(geo).__mro__]
            if "GeodesicSpace" not in mro:
                raise TypeError("cannot translate geodesic %s" % type(geo).__name__)
            spec = MetricSpec.euclidean(self._dev.spec.dim)
        self._geo_cache[key] = (w, spec)
        return spec

    def _generic(self, interpolator):
        l = interpolator.length()
        k = int(math.ceil(l / self.resolution))
        pts = [interpolator.start(), interpolator.end()]
        for i in range(k):
            pts.append(interpolator.eval(float(i + 1) / float(k + 2)))
        return bool(self._dev._be.feasible(np.asarray(pts, dtype=np.float64)).all())


class DeviceControlSpace(object):
    """GPU nextState for a pomp ControlSpace; everything else is forwarded to the wrapped space."""

    def __init__(self, control_space, _backend_factory=None):
        self.base = control_space
        self.spec = dynamics_from_pomp(control_space)
        self._be = (_backend_factory or _dyn_backend)(self.spec)

    def __getattr__(self, name):  # configurationSpace, controlSet, interpolator, connection, ...
        return getattr(self.base, name)

    def nextState(self, x, u):
        return self._be.next_state([x], [u])[0].tolist()

    def nextState_batch(self, xs, us):
        return self._be.next_state(xs, us)


class BatchedRandomControlSelector(object):
    """RandomControlSelector with the numSamples candidates evaluated in one fused launch.

    Controls are drawn on the host with the control set's own sample() in the reference's order,
    so Python's `random` stream -- and with it the planner's tree -- is unchanged."""

    def __init__(self, controlSpace, metric, numSamples, _backend_factory=None):
        self.controlSpace = controlSpace
        self.metric = metric
        self.numSamples = numSamples
        base = controlSpace.base if isinstance(controlSpace, DeviceControlSpace) else controlSpace
        self._dyn = dynamics_from_pomp(base)
        self._be = (_backend_factory or _dyn_backend)(self._dyn)
        self._mspec = None
        self._msig = None

    def _metric_spec(self):
        sig = (getattr(self.metric, "costWeight", None), getattr(self.metric, "costMax", None))
        if self._mspec is None or sig != self._msig:
            self._mspec = metric_from_pomp(self.metric, self._dyn.state_dim())
            self._msig = sig
        return self._mspec

    def select(self, x, xdesired):
        U = self.controlSpace.controlSet(x)
        if self.numSamples == 1:
            u = U.sample()
            return u if U.contains(u) else None
        cands = [U.sample() for _ in range(self.numSamples)]
        valid = [1 if U.contains(u) else 0 for u in cands]
        best, _, _ = self._be.select_control(self._metric_spec(), [x], [xdesired], [cands], [valid])
        b = int(best[0])
        return None if b < 0 else cands[b]

    def select_batch(self, xs, xdesired, controls, valid=None):
        """xs[B], xdesired[B], controls[B][S][u]; returns (best[B], xnext_best[B], dist_best[B])."""
        return self._be.select_control(self._metric_spec(), xs, xdesired, controls, valid)
