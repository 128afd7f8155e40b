"""ctypes front end of oracle/libpomp_oracle.so -- TEST INFRASTRUCTURE ONLY.

The C file restates the reference algorithms (each function cites pomp file:line); this module
only marshals numpy arrays.  Allowed importers: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from pyoptimalmotionplanning_b200 import _abi
from pyoptimalmotionplanning_b200.descriptors import MetricSpec, SpaceSpec, DynamicsSpec  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpomp_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "pomp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_metric.restype = C.c_double
        _lib.orc_nn_neighbors.restype = C.c_int64
        _lib.orc_kd_create.restype = C.c_void_p
        _lib.orc_kd_neighbors.restype = C.c_int64
        _lib.orc_kd_stats.restype = C.c_int64
    return _lib


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _mask(mask):
    return None if mask is None else np.ascontiguousarray(np.asarray(mask, dtype=np.uint8))


def metric(m, a, b):
    a, b = _f64(a), _f64(b)
    return lib().orc_metric(C.byref(m.desc()), _p(a), _p(b))


def nn_nearest(m, pts, q, mask=None):
    pts, q, mask = _f64(pts).reshape(-1, m.dim), _f64(q).reshape(-1, m.dim), _mask(mask)
    idx = np.empty(len(q), np.int64)
    dist = np.empty(len(q), np.float64)
    lib().orc_nn_nearest(C.byref(m.desc()), _p(pts), C.c_int64(len(pts)), _p(mask), _p(q),
                         C.c_int64(len(q)), _p(idx), _p(dist))
    return idx, dist


def nn_knearest(m, pts, q, k, mask=None):
    pts, q, mask = _f64(pts).reshape(-1, m.dim), _f64(q).reshape(-1, m.dim), _mask(mask)
    idx = np.empty((len(q), k), np.int64)
    dist = np.empty((len(q), k), np.float64)
    cnt = np.empty(len(q), np.int32)
    lib().orc_nn_knearest(C.byref(m.desc()), _p(pts), C.c_int64(len(pts)), _p(mask), _p(q),
                          C.c_int64(len(q)), C.c_int(k), _p(idx), _p(dist), _p(cnt))
    return idx, dist, cnt


def nn_neighbors(m, pts, q, radius, inclusive=True, mask=None):
    pts, q, mask = _f64(pts).reshape(-1, m.dim), _f64(q).reshape(-1, m.dim), _mask(mask)
    off = np.empty(len(q) + 1, np.int64)
    cap = 1 << 16
    while True:
        idx = np.empty(cap, np.int64)
        need = lib().orc_nn_neighbors(C.byref(m.desc()), _p(pts), C.c_int64(len(pts)), _p(mask), _p(q),
                                      C.c_int64(len(q)), C.c_double(radius), C.c_int(1 if inclusive else 0),
                                      _p(off), _p(idx), C.c_int64(cap))
        if need <= cap:
            return off, idx[:need]
        cap = int(need)


class KDTree(object):
    """The reference's default index (pomp/structures/kdtree.py) restated in C."""

    def __init__(self, m, pts=None):
        self.m = m
        self.desc = m.desc()
        self.pts = None
        self.h = C.c_void_p(lib().orc_kd_create(C.byref(self.desc), None))
        if pts is not None:
            self.set(pts)

    def set(self, pts):
        self.pts = _f64(pts).reshape(-1, self.m.dim)
        lib().orc_kd_set(self.h, _p(self.pts), C.c_int64(len(self.pts)))

    def add_all_incrementally(self, pts):
        """Insert one by one through KDTree.add, as NearestNeighbors.add does."""
        self.pts = _f64(pts).reshape(-1, self.m.dim)
        for i in range(len(self.pts)):
            lib().orc_kd_add(self.h, _p(self.pts), C.c_int64(i))

    def nearest(self, q, mask=None):
        q, mask = _f64(q).reshape(-1, self.m.dim), _mask(mask)
        idx = np.empty(len(q), np.int64)
        dist = np.empty(len(q), np.float64)
        lib().orc_kd_nearest(self.h, _p(mask), _p(q), C.c_int64(len(q)), _p(idx), _p(dist))
        return idx, dist

    def knearest(self, q, k, mask=None):
        q, mask = _f64(q).reshape(-1, self.m.dim), _mask(mask)
        idx = np.empty((len(q), k), np.int64)
        dist = np.empty((len(q), k), np.float64)
        cnt = np.empty(len(q), np.int32)
        lib().orc_kd_knearest(self.h, _p(mask), _p(q), C.c_int64(len(q)), C.c_int(k), _p(idx), _p(dist), _p(cnt))
        return idx, dist, cnt

    def neighbors(self, q, rad):
        q = _f64(q).reshape(self.m.dim)
        cap = 1024
        while True:
            out = np.empty(cap, np.int64)
            n = lib().orc_kd_neighbors(self.h, _p(q), C.c_double(rad), _p(out), C.c_int64(cap))
            if n <= cap:
                return out[:n]
            cap = int(n)

    def stats(self):
        return {"num_nodes": lib().orc_kd_stats(self.h, 0), "max_depth": lib().orc_kd_stats(self.h, 1)}

    def __del__(self):
        try:
            lib().orc_kd_destroy(self.h)
        except Exception:
            pass


def feasible(s, x):
    x = _f64(x).reshape(-1, s.dim)
    ok = np.empty(len(x), np.uint8)
    d = s.desc()
    lib().orc_feasible(C.byref(d), _p(x), C.c_int64(len(x)), _p(ok))
    return ok.astype(bool)


def edge_check(s, a, b, res):
    a, b = _f64(a).reshape(-1, s.dim), _f64(b).reshape(-1, s.dim)
    ok = np.empty(len(a), np.uint8)
    d = s.desc()
    lib().orc_edge_check(C.byref(d), _p(a), _p(b), C.c_int64(len(a)), C.c_double(res), _p(ok))
    return ok.astype(bool)


def edge_check_indexed(s, nodes, ij, res):
    nodes = _f64(nodes).reshape(-1, s.dim)
    ij = np.ascontiguousarray(np.asarray(ij, dtype=np.int32).reshape(-1, 2))
    ok = np.empty(len(ij), np.uint8)
    d = s.desc()
    lib().orc_edge_check_indexed(C.byref(d), _p(nodes), _p(ij), C.c_int64(len(ij)), C.c_double(res), _p(ok))
    return ok.astype(bool)


def edge_check_path(s, g, waypoints, offsets, res):
    wp = _f64(waypoints).reshape(-1, s.dim)
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32))
    ok = np.empty(len(off) - 1, np.uint8)
    d = s.desc()
    lib().orc_edge_check_path(C.byref(d), C.byref(g.desc()), _p(wp), _p(off), C.c_int64(len(off) - 1),
                              C.c_double(res), _p(ok))
    return ok.astype(bool)


def edge_check_control(s, dyn, g, x, u, res):
    x, u = _f64(x).reshape(-1, dyn.x_dim), _f64(u).reshape(-1, dyn.u_dim)
    ok = np.empty(len(x), np.uint8)
    d = s.desc()
    lib().orc_edge_check_control(C.byref(d), C.byref(dyn.desc()), C.byref(g.desc()), _p(x), _p(u),
                                 C.c_int64(len(x)), C.c_double(res), _p(ok))
    return ok.astype(bool)


def next_state(dyn, x, u):
    X = dyn.state_dim()
    x, u = _f64(x).reshape(-1, X), _f64(u).reshape(-1, dyn.u_dim)
    out = np.empty_like(x)
    lib().orc_next_state(C.byref(dyn.desc()), _p(x), _p(u), C.c_int64(len(x)), _p(out))
    return out


def select_control(dyn, m, x, xdes, u, valid=None):
    X = dyn.state_dim()
    x, xdes = _f64(x).reshape(-1, X), _f64(xdes).reshape(-1, X)
    B = len(x)
    u = _f64(u).reshape(B, -1, dyn.u_dim)
    S = u.shape[1]
    valid = None if valid is None else np.ascontiguousarray(np.asarray(valid, dtype=np.uint8).reshape(B, S))
    best = np.empty(B, np.int32)
    xb = np.empty((B, X), np.float64)
    db = np.empty(B, np.float64)
    lib().orc_select_control(C.byref(dyn.desc()), C.byref(m.desc()), _p(x), _p(xdes), _p(u), _p(valid),
                             C.c_int64(B), C.c_int(S), _p(best), _p(xb), _p(db))
    return best, xb, db


def rrt_extend(m, pts, s, xrand, max_step, res, mask=None):
    pts, xrand, mask = _f64(pts).reshape(-1, m.dim), _f64(xrand).reshape(m.dim), _mask(mask)
    parent = C.c_int64(-1)
    added = C.c_int32(0)
    xnew = np.zeros(m.dim, np.float64)
    d = s.desc()
    lib().orc_rrt_extend(C.byref(m.desc()), _p(pts), C.c_int64(len(pts)), _p(mask), C.byref(d), _p(xrand),
                         C.c_double(max_step), C.c_double(res), C.byref(parent), _p(xnew), C.byref(added))
    return parent.value, xnew, bool(added.value)


def nn_density(metric_spec, pts, q, radius, sigma, inclusive=True, mask=None):
    """ESTKinodynamicPlanner.density (pomp/planners/kinodynamicplanner.py:688-693): neighbors(x, radius), then
    d += math.exp(-(dist/sigma)**2) over the hits, here in insertion order.  Returns (density[nq], count[nq])."""
    import math
    pts = np.asarray(pts, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64).reshape(-1, pts.shape[1])
    off, hits = nn_neighbors(metric_spec, pts, q, radius, inclusive, mask)
    dens = np.zeros(len(q))
    cnt = np.zeros(len(q), dtype=np.int32)
    for i in range(len(q)):
        d = 0.0
        for j in hits[off[i]:off[i + 1]]:
            dist = float(metric(metric_spec, pts[j], q[i]))
            d += math.exp(-(dist / sigma) ** 2)
        dens[i] = d
        cnt[i] = off[i + 1] - off[i]
    return dens, cnt


def set_threads(n):
    """OpenMP threads used by the batched oracle loops; returns the count in effect."""
    L = lib()
    L.orc_set_threads.restype = C.c_int
    L.orc_set_threads.argtypes = [C.c_int]
    return int(L.orc_set_threads(int(n)))


def est_projection_density(bases, pts, q):
    """ESTWithProjections.density (kinodynamicplanner.py:834-856) restated with dictionaries, as the reference keeps
    them (:817-832).  bases: [(indices, scales, weight)]; pts: the nodes; q: query states."""
    tables = []
    for idx, sc, w in bases:
        tab = {}
        for p_ in pts:
            key = tuple(int(0.0 + float(p_[k]) * v) for k, v in zip(idx, sc))
            tab[key] = tab.get(key, 0) + 1
        tables.append(tab)
    out = []
    for x in q:
        c = 0.0
        for (idx, sc, w), tab in zip(bases, tables):
            key = tuple(int(0.0 + float(x[k]) * v) for k, v in zip(idx, sc))
            c += float(tab.get(key, 0)) * w
        out.append(c / len(bases))
    return np.array(out)
