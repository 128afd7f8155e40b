"""Shared helpers: load tests/golden fixtures and turn their JSON specs into descriptors."""
import json
import os

from pyoptimalmotionplanning_b200 import _abi
from pyoptimalmotionplanning_b200.descriptors import MetricSpec, SpaceSpec, DynamicsSpec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INF = float("inf")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def metric_spec(s):
    k = s["kind"]
    if k == "euclidean":
        m = MetricSpec.euclidean(s["dim"])
    elif k == "l1":
        m = MetricSpec.l1(s["dim"])
    elif k == "linf":
        m = MetricSpec.linf(s["dim"])
    elif k == "weighted":
        m = MetricSpec.weighted_euclidean(s["w"])
    elif k == "scaled":
        m = MetricSpec.scaled_euclidean(s["dim"], s["w"])
    elif k == "product":
        m = MetricSpec.geodesic_product([tuple(c) for c in s["comps"]], s["w"])
    else:
        raise ValueError(k)
    if "cost" in s:
        m = m.with_cost(s["cost"]["weight"], s["cost"]["max"])
    return m


def space_spec(s, obstacle_dims=(0, 1)):
    lo = [-INF if v is None else v for v in s["lo"]]
    hi = [INF if v is None else v for v in s["hi"]]
    return SpaceSpec(lo, hi, s.get("boxes", []), s.get("circles", []), obstacle_dims)


_DYN = {"adaptor": _abi.DYN_ADAPTOR, "cv": _abi.DYN_CV, "cv_euler": _abi.DYN_CV_EULER, "dubins": _abi.DYN_DUBINS,
        "pendulum": _abi.DYN_PENDULUM, "flappy": _abi.DYN_FLAPPY}
_COST = {None: _abi.COST_NONE, "time": _abi.COST_TIME, "abs_u0": _abi.COST_ABS_U0,
         "path_length": _abi.COST_PATH_LENGTH}


def dyn_spec(s):
    k = s["kind"]
    cost = _COST[s.get("cost")]
    if k == "dubins":
        return DynamicsSpec(_abi.DYN_DUBINS, 3, 2, cost)
    if k == "pendulum":
        return DynamicsSpec(_abi.DYN_PENDULUM, 2, 2, cost, dt=s["dt"], p=s["p"])
    if k in ("cv", "cv_euler"):
        return DynamicsSpec(_DYN[k], s["x_dim"], s["x_dim"] // 2 + 1, cost, dt=s["dt"])
    if k == "flappy":
        return DynamicsSpec(_abi.DYN_FLAPPY, 3, 2, cost, p=s["p"])
    if k == "adaptor":
        return DynamicsSpec(_abi.DYN_ADAPTOR, s["x_dim"], s["x_dim"], cost)
    if k == "lti":
        import numpy as np
        A, B = np.asarray(s["A"], dtype=np.float64), np.asarray(s["B"], dtype=np.float64)
        return DynamicsSpec(_abi.DYN_LTI, A.shape[0], B.shape[1], cost, A=A, B=B)
    raise ValueError(k)
