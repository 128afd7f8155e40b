"""Host-side descriptors of metrics, feasibility spaces and dynamics, and the translators that
derive them from (unmodified) pomp objects by duck typing.

Nothing here imports pomp: the reference objects are recognised by class name and attributes,
so the same code serves pomp's own classes and the light stand-ins used in the GPU tests.
Unknown objects raise TypeError -- there is no CPU fallback to hand them to.

Reference shapes mirrored here:
  metrics    pomp/spaces/metric.py:4-23, pomp/spaces/geodesicspace.py:28-52,
             pomp/planners/kinodynamicplanner.py:899-913 (CostMetric)
  spaces     pomp/spaces/configurationspace.py:127-199, pomp/spaces/sets.py:166-182,
             pomp/example_problems/geometric.py:10-60
  dynamics   pomp/spaces/controlspace.py:66-271, statespace.py:41-70, costspace.py:10-41,
             example_problems/dubins.py:90-125, pendulum.py:116-154, flappy.py:10-31
"""
import math

import numpy as np

from . import _abi

INF = float("inf")


# ---------------------------------------------------------------------------- metrics
class MetricSpec(object):
    """Plain description of a pomp metric callable; `.desc()` gives the C struct."""

    def __init__(self, kind, dim, comp_kind=(), comp_dim=(), comp_weight=(), dim_weight=(), scale=1.0,
                 has_cost=False, cost_weight=None, cost_max=None):
        self.kind = kind
        self.dim = int(dim)
        self.comp_kind = list(comp_kind)
        self.comp_dim = list(comp_dim)
        self.comp_weight = [float(w) for w in comp_weight]
        self.dim_weight = [float(w) for w in dim_weight]
        self.scale = float(scale)
        self.has_cost = bool(has_cost)
        self.cost_weight = cost_weight
        self.cost_max = cost_max
        if self.dim < 1 or self.dim > _abi.MAX_DIM:
            raise ValueError("metric dimension %d outside 1..%d" % (self.dim, _abi.MAX_DIM))
        if len(self.comp_kind) > _abi.MAX_COMP:
            raise ValueError("too many product components")

    # constructors -------------------------------------------------------------
    @staticmethod
    def euclidean(dim):
        return MetricSpec(_abi.METRIC_EUCLIDEAN, dim)

    @staticmethod
    def l1(dim):
        return MetricSpec(_abi.METRIC_L1, dim)

    @staticmethod
    def linf(dim):
        return MetricSpec(_abi.METRIC_LINF, dim)

    @staticmethod
    def weighted_euclidean(w):
        if hasattr(w, "__iter__"):
            w = list(w)
            return MetricSpec(_abi.METRIC_WEIGHTED_EUCLIDEAN, len(w), dim_weight=w)
        raise TypeError("scalar weight needs a dimension: use MetricSpec.scaled_euclidean(dim, w)")

    @staticmethod
    def scaled_euclidean(dim, w):
        return MetricSpec(_abi.METRIC_SCALED_EUCLIDEAN, dim, scale=w)

    @staticmethod
    def geodesic_product(components, weights=None):
        """components: list of ('R', d) / ('SO2',) entries, weights per component."""
        kinds, dims = [], []
        for c in components:
            if c[0] == "SO2":
                kinds.append(_abi.COMP_SO2)
                dims.append(1)
            else:
                kinds.append(_abi.COMP_CARTESIAN)
                dims.append(int(c[1]))
        if weights is None:
            weights = [1.0] * len(kinds)
        return MetricSpec(_abi.METRIC_GEODESIC_PRODUCT, sum(dims), kinds, dims, weights)

    def with_cost(self, cost_weight=1.0, cost_max=None):
        """CostMetric(base, costWeight) over states carrying a trailing cost coordinate."""
        m = MetricSpec(self.kind, self.dim + 1, self.comp_kind, self.comp_dim, self.comp_weight,
                       self.dim_weight, self.scale, True, cost_weight, cost_max)
        return m

    def base_dim(self):
        return self.dim - 1 if self.has_cost else self.dim

    def desc(self):
        d = _abi.MetricDesc()
        d.kind = self.kind
        d.dim = self.dim
        d.n_comp = len(self.comp_kind)
        for i, (k, n, w) in enumerate(zip(self.comp_kind, self.comp_dim, self.comp_weight)):
            d.comp_kind[i] = k
            d.comp_dim[i] = n
            d.comp_weight[i] = w
        for i, w in enumerate(self.dim_weight):
            d.dim_weight[i] = w
        d.scale = self.scale
        d.has_cost = 1 if self.has_cost else 0
        d.cost_weight_set = 1 if (self.has_cost and self.cost_weight is not None) else 0
        d.cost_max_set = 1 if (self.has_cost and self.cost_max is not None) else 0
        d.cost_weight = float(self.cost_weight) if self.cost_weight is not None else 0.0
        d.cost_max = float(self.cost_max) if self.cost_max is not None else INF
        return d

    def key(self):
        return (self.kind, self.dim, tuple(self.comp_kind), tuple(self.comp_dim), tuple(self.comp_weight),
                tuple(self.dim_weight), self.scale, self.has_cost, self.cost_weight, self.cost_max)


def _geodesic_components(geodesic):
    """MultiGeodesicSpace -> (components, weights) (geodesicspace.py:28-52)."""
    comps = []
    for c in geodesic.components:
        name = type(c).__name__
        mro = [k.__name__ for k in type(c).__mro__]
        if "SO2Geodesic" in mro or "SO2Space" in mro:
            comps.append(("SO2",))
        elif "MultiGeodesicSpace" in mro:
            raise TypeError("nested product geodesics are not supported (%s)" % name)
        elif "GeodesicSpace" in mro or hasattr(c, "dimension"):
            # any plain GeodesicSpace measures vectorops.distance (geodesicspace.py:7-8)
            if type(c).distance is not _find_base(type(c), "GeodesicSpace").distance:
                raise TypeError("component %s overrides distance(); cannot translate" % name)
            comps.append(("R", c.dimension()))
        else:
            raise TypeError("cannot translate geodesic component %s" % name)
    return comps, list(geodesic.componentWeights)


def _find_base(cls, name):
    for k in cls.__mro__:
        if k.__name__ == name:
            return k
    raise TypeError("%s does not derive from %s" % (cls.__name__, name))


def metric_from_pomp(metric, dim=None):
    """Translate a pomp metric callable (or a MetricSpec) into a MetricSpec.

    `dim` is needed for the dimension-free callables (euclideanMetric, L1Metric, ...); the
    NearestNeighbors drop-in supplies it from the first point it sees."""
    if isinstance(metric, MetricSpec):
        return metric
    name = getattr(metric, "__name__", type(metric).__name__)
    tname = type(metric).__name__
    if tname == "CostMetric":  # kinodynamicplanner.py:899-913
        base = metric_from_pomp(metric.metric, None if dim is None else dim - 1)
        return base.with_cost(metric.costWeight, metric.costMax)
    if tname == "WeightedEuclideanMetric":  # metric.py:16-23
        if hasattr(metric.w, "__iter__"):
            return MetricSpec.weighted_euclidean(metric.w)
        if dim is None:
            raise TypeError("scalar WeightedEuclideanMetric needs a dimension")
        return MetricSpec.scaled_euclidean(dim, metric.w)
    if tname in ("function", "builtin_function_or_method"):
        if name in ("euclideanMetric", "L2Metric", "distance"):  # metric.py:4-8, vectorops.distance
            mod = getattr(metric, "__module__", "") or ""
            if name == "distance" and not mod.endswith("vectorops"):
                raise TypeError("unknown distance function %s.%s" % (mod, name))
            if dim is None:
                raise TypeError("euclidean metric needs a dimension")
            return MetricSpec.euclidean(dim)
        if name == "L1Metric":
            return MetricSpec.l1(dim)
        if name == "LinfMetric":
            return MetricSpec.linf(dim)
    if tname == "method" and name == "distance":
        owner = metric.__self__
        # MultiConfigurationSpace.distance -> self.geodesic.distance (configurationspace.py:201-202)
        geo = getattr(owner, "geodesic", owner)
        if hasattr(geo, "components") and hasattr(geo, "componentWeights"):
            comps, w = _geodesic_components(geo)
            return MetricSpec.geodesic_product(comps, w)
        mro = [k.__name__ for k in type(owner).__mro__]
        if "GeodesicSpace" in mro or "ConfigurationSpace" in mro:
            # GeodesicSpace.distance / ConfigurationSpace.distance are both Euclidean
            d = dim if dim is not None else owner.dimension()
            return MetricSpec.euclidean(d)
    raise TypeError("cannot translate metric %r to a device descriptor (no CPU fallback)" % (metric,))


# ---------------------------------------------------------------------------- spaces
class SpaceSpec(object):
    """Box-bounded product space with 2-D box/disc obstacles on two of its coordinates."""

    def __init__(self, lo, hi, boxes=(), circles=(), obstacle_dims=(0, 1)):
        self.lo = [float(v) for v in lo]
        self.hi = [float(v) for v in hi]
        assert len(self.lo) == len(self.hi)
        self.dim = len(self.lo)
        if self.dim < 1 or self.dim > _abi.MAX_DIM:
            raise ValueError("space dimension %d outside 1..%d" % (self.dim, _abi.MAX_DIM))
        self.boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64).reshape(-1, 4))
        self.circles = np.ascontiguousarray(np.asarray(circles, dtype=np.float64).reshape(-1, 3))
        self.obstacle_dims = (int(obstacle_dims[0]), int(obstacle_dims[1]))
        if len(self.boxes) or len(self.circles):
            assert max(self.obstacle_dims) < self.dim

    def desc(self):
        d = _abi.SpaceDesc()
        d.dim = self.dim
        d.obstacle_dim0, d.obstacle_dim1 = self.obstacle_dims
        d.n_boxes = len(self.boxes)
        d.n_circles = len(self.circles)
        for i in range(self.dim):
            d.lo[i] = self.lo[i]
            d.hi[i] = self.hi[i]
        d.boxes_h = self.boxes.ctypes.data if len(self.boxes) else None
        d.circles_h = self.circles.ctypes.data if len(self.circles) else None
        d._keepalive = (self.boxes, self.circles)
        return d


def _component_bounds(c):
    """(lo, hi, boxes, circles) of one ConfigurationSpace component, by duck typing."""
    mro = [k.__name__ for k in type(c).__mro__]
    boxes, circles = [], []
    if "Geometric2DCSpace" in mro:  # geometric.py:48-60
        for o in c.obstacles:
            omro = [k.__name__ for k in type(o).__mro__]
            if "BoxSet" in omro:
                boxes.append([o.bmin[0], o.bmin[1], o.bmax[0], o.bmax[1]])
            elif "Circle" in omro:
                circles.append([o.center[0], o.center[1], o.radius])
            else:
                raise TypeError("cannot translate obstacle %s" % type(o).__name__)
        return list(c.box.bmin), list(c.box.bmax), boxes, circles
    if "BoxConfigurationSpace" in mro:  # configurationspace.py:127-143
        return list(c.box.bmin), list(c.box.bmax), boxes, circles
    if "MultiConfigurationSpace" in mro:
        raise TypeError("nested MultiConfigurationSpace is not supported")
    if "ConfigurationSpace" in mro:
        # SO2Space, CartesianConfigurationSpace, GeodesicConfigurationSpace inherit
        # ConfigurationSpace.feasible -> True (configurationspace.py:46-48)
        base = _find_base(type(c), "ConfigurationSpace")
        if type(c).feasible is not base.feasible:
            raise TypeError("space %s overrides feasible(); cannot translate" % type(c).__name__)
        d = c.dimension()
        return [-INF] * d, [INF] * d, boxes, circles
    raise TypeError("cannot translate space component %s" % type(c).__name__)


def space_from_pomp(space):
    """Translate a pomp ConfigurationSpace into a SpaceSpec (feasible() semantics only)."""
    if isinstance(space, SpaceSpec):
        return space
    mro = [k.__name__ for k in type(space).__mro__]
    if "MultiConfigurationSpace" in mro:  # configurationspace.py:165-199
        lo, hi, boxes, circles, odims = [], [], [], [], (0, 1)
        for c in space.components:
            l, h, b, ci = _component_bounds(c)
            if b or ci:
                if boxes or circles:
                    raise TypeError("only one obstacle-carrying component is supported")
                odims = (len(lo), len(lo) + 1)
                boxes, circles = b, ci
            lo += l
            hi += h
        return SpaceSpec(lo, hi, boxes, circles, odims)
    l, h, b, ci = _component_bounds(space)
    return SpaceSpec(l, h, b, ci, (0, 1))


# ---------------------------------------------------------------------------- dynamics
class DynamicsSpec(object):
    def __init__(self, kind, x_dim, u_dim, cost_kind=_abi.COST_NONE, dt=0.0, p=(), A=None, B=None):
        self.kind = kind
        self.x_dim = int(x_dim)
        self.u_dim = int(u_dim)
        self.cost_kind = cost_kind
        self.dt = float(dt)
        self.p = [float(v) for v in p]
        self.A = None if A is None else np.asarray(A, dtype=np.float64)
        self.B = None if B is None else np.asarray(B, dtype=np.float64)
        if self.x_dim + (1 if cost_kind else 0) > _abi.MAX_DIM or self.u_dim > _abi.MAX_DIM:
            raise ValueError("dynamics dimensions exceed POMP_MAX_DIM")

    def state_dim(self):
        return self.x_dim + (1 if self.cost_kind != _abi.COST_NONE else 0)

    def with_cost(self, cost_kind):
        return DynamicsSpec(self.kind, self.x_dim, self.u_dim, cost_kind, self.dt, self.p, self.A, self.B)

    def desc(self):
        d = _abi.DynDesc()
        d.kind, d.x_dim, d.u_dim, d.cost_kind, d.dt = self.kind, self.x_dim, self.u_dim, self.cost_kind, self.dt
        for i, v in enumerate(self.p):
            d.p[i] = v
        if self.A is not None:
            for i, v in enumerate(self.A.reshape(-1)):
                d.A[i] = v
        if self.B is not None:
            for i, v in enumerate(self.B.reshape(-1)):
                d.B[i] = v
        return d


def _objective_cost_kind(objective, base_kind):
    n = type(objective).__name__
    if n == "TimeObjectiveFunction":
        return _abi.COST_TIME
    if n == "DubinsCarDistanceObjectiveFunction" and getattr(objective, "n", 1) == 1:
        return _abi.COST_ABS_U0
    if n == "PathLengthObjectiveFunction" and base_kind == _abi.DYN_ADAPTOR:
        return _abi.COST_PATH_LENGTH
    raise TypeError("cannot translate objective %s" % n)


def dynamics_from_pomp(cs, stepwise=False):
    """Translate a pomp ControlSpace into a DynamicsSpec.

    stepwise=True selects the integrated trajectory()[-1] variant where the reference has both a
    closed form nextState and a stepwise interpolator (CVControlSpace, statespace.py:47-70)."""
    if isinstance(cs, DynamicsSpec):
        return cs
    mro = [k.__name__ for k in type(cs).__mro__]
    if "CostControlSpace" in mro:  # costspace.py:10-41
        base = dynamics_from_pomp(cs.baseSpace, stepwise)
        return base.with_cost(_objective_cost_kind(cs.objective, base.kind))
    if "ControlSpaceAdaptor" in mro:  # controlspace.py:66-86
        d = cs.cspace.dimension()
        return DynamicsSpec(_abi.DYN_ADAPTOR, d, d)
    if "CVControlSpace" in mro:  # statespace.py:41-70
        d = cs.cspace.dimension()
        return DynamicsSpec(_abi.DYN_CV_EULER if stepwise else _abi.DYN_CV, d, d // 2 + 1, dt=cs.dt)
    if "DubinsCarSpace" in mro:  # dubins.py:90-125
        return DynamicsSpec(_abi.DYN_DUBINS, 3, 2)
    if "LambdaKinodynamicSpace" in mro:  # controlspace.py:224-271
        owner = getattr(cs.f, "__self__", None)
        if owner is not None and type(owner).__name__ == "Pendulum" and cs.f.__name__ == "derivative":
            return DynamicsSpec(_abi.DYN_PENDULUM, 2, 2, dt=cs.dt, p=[owner.g, owner.m, owner.L])
        raise TypeError("cannot translate the dynamics function of %r" % (cs,))
    if "LTIControlSpace" in mro:  # controlspace.py:88-112
        A, B = np.asarray(cs.A, dtype=np.float64), np.asarray(cs.B, dtype=np.float64)
        return DynamicsSpec(_abi.DYN_LTI, A.shape[0], B.shape[1], A=A, B=B)
    if "FlappyControlSpace" in mro:  # flappy.py:10-31
        f = cs.flappy
        return DynamicsSpec(_abi.DYN_FLAPPY, 3, 2, p=[f.v_x, f.gravity, f.thrust])
    raise TypeError("cannot translate control space %s (no CPU fallback)" % type(cs).__name__)


def geodesic_of_space(space, dim):
    """The geodesic a PiecewiseLinearInterpolator on `space` uses (controlspace.py:169-172)."""
    geo = getattr(space, "geodesic", None)
    if geo is not None and hasattr(geo, "components"):
        comps, w = _geodesic_components(geo)
        return MetricSpec.geodesic_product(comps, w)
    return MetricSpec.euclidean(dim)
