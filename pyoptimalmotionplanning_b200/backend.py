"""Thin object layer over the C-ABI (csrc/libpomp_b200.so): owns handles, marshals numpy/torch
buffers into plain pointers.  This is the ONLY compute backend of the product -- every method
ends in a CUDA kernel launch; without the library or a device it raises.

(The classes here define the small internal interface that the host-side drop-ins use; the CPU
test-suite substitutes an oracle-backed fake with the same methods to exercise the host logic of
the drop-ins where no GPU exists.  That fake lives under tests/, never here.)
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import check


def _f64(a, cols=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _is_torch_cuda(t):
    return hasattr(t, "is_cuda") and t.is_cuda


class CudaNN(object):
    """NearestNeighbors state on one GPU (pomp_nn_* entry points)."""

    def __init__(self, metric_spec, device=0, capacity_hint=0):
        self.lib = _abi.load_library()
        self.spec = metric_spec
        self.dim = metric_spec.dim
        self.device = device
        self._h = C.c_void_p()
        check(self.lib, self.lib.pomp_nn_create(C.byref(metric_spec.desc()), capacity_hint, device, C.byref(self._h)))
        # reusable single-query buffers
        self._q1 = (C.c_double * self.dim)()
        self._i1 = C.c_int64()
        self._d1 = C.c_double()
        self._first = C.c_int64()

    def close(self):
        if self._h is not None and self._h.value:
            self.lib.pomp_nn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- mutation
    def reset(self):
        check(self.lib, self.lib.pomp_nn_reset(self._h))

    def size(self):
        return self.lib.pomp_nn_size(self._h)

    def add(self, pts):
        pts = _f64(pts, self.dim)
        check(self.lib, self.lib.pomp_nn_add_h(self._h, _ptr(pts), len(pts), C.byref(self._first)))
        return self._first.value

    def add_one(self, pt):
        q = self._q1
        for j in range(self.dim):
            q[j] = pt[j]
        check(self.lib, self.lib.pomp_nn_add_h(self._h, q, 1, C.byref(self._first)))
        return self._first.value

    def set(self, pts):
        pts = _f64(pts, self.dim)
        check(self.lib, self.lib.pomp_nn_set_h(self._h, _ptr(pts), len(pts)))

    def set_device(self, ptr, n, stream=None):
        check(self.lib, self.lib.pomp_nn_set(self._h, C.c_void_p(ptr), n, C.c_void_p(stream or 0)))

    def remove(self, index):
        check(self.lib, self.lib.pomp_nn_remove(self._h, index))

    def set_mask(self, mask):
        if mask is None:
            check(self.lib, self.lib.pomp_nn_set_mask_h(self._h, None, 0))
        else:
            m = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8))
            check(self.lib, self.lib.pomp_nn_set_mask_h(self._h, _ptr(m), len(m)))

    def update_metric(self, metric_spec):
        self.spec = metric_spec
        check(self.lib, self.lib.pomp_nn_update_metric(self._h, C.byref(metric_spec.desc())))

    def set_param(self, name, value):
        check(self.lib, self.lib.pomp_nn_set_param(self._h, name.encode(), float(value)))

    def get_stat(self, name):
        out = C.c_double()
        check(self.lib, self.lib.pomp_nn_get_stat(self._h, name.encode(), C.byref(out)))
        return out.value

    def build_index(self, stream=None):
        check(self.lib, self.lib.pomp_nn_build_index(self._h, C.c_void_p(stream or 0)))

    # ---- queries, host buffers
    def nearest_one(self, pt):
        q = self._q1
        for j in range(self.dim):
            q[j] = pt[j]
        check(self.lib, self.lib.pomp_nn_nearest_h(self._h, q, 1, C.byref(self._i1), C.byref(self._d1)))
        return self._i1.value, self._d1.value

    def nearest(self, q):
        q = _f64(q, self.dim)
        idx = np.empty(len(q), np.int64)
        dist = np.empty(len(q), np.float64)
        check(self.lib, self.lib.pomp_nn_nearest_h(self._h, _ptr(q), len(q), _ptr(idx), _ptr(dist)))
        return idx, dist

    def knearest(self, q, k):
        q = _f64(q, self.dim)
        idx = np.empty((len(q), k), np.int64)
        dist = np.empty((len(q), k), np.float64)
        cnt = np.empty(len(q), np.int32)
        check(self.lib, self.lib.pomp_nn_knearest_h(self._h, _ptr(q), len(q), k, _ptr(idx), _ptr(dist), _ptr(cnt)))
        return idx, dist, cnt

    def neighbors(self, q, radius, inclusive=True):
        q = _f64(q, self.dim)
        off = np.empty(len(q) + 1, np.int64)
        cap = max(64, 32 * len(q))
        need = C.c_int64()
        while True:
            idx = np.empty(cap, np.int64)
            rc = self.lib.pomp_nn_neighbors_h(self._h, _ptr(q), len(q), float(radius), 1 if inclusive else 0,
                                              _ptr(off), _ptr(idx), cap, C.byref(need))
            if rc == _abi.POMP_ERR_CAPACITY:
                cap = int(need.value)
                continue
            check(self.lib, rc)
            return off, idx[:need.value]

    # ---- queries, device buffers (raw pointers: torch tensors' data_ptr())
    def density(self, q, radius, sigma, inclusive=True):
        """pomp_nn_density_h: (density[nq], count[nq]) -- Gaussian kernel density of EST."""
        q = _f64(q, self.dim)
        nq = len(q)
        dens = np.empty(nq, dtype=np.float64)
        cnt = np.empty(nq, dtype=np.int32)
        check(self.lib, self.lib.pomp_nn_density_h(self._h, _ptr(q), nq, float(radius), float(sigma),
                                                   1 if inclusive else 0, _ptr(dens), _ptr(cnt)))
        return dens, cnt

    def density_device(self, q_ptr, nq, radius, sigma, inclusive, dens_ptr, cnt_ptr=None, stream=None):
        check(self.lib, self.lib.pomp_nn_density(self._h, C.c_void_p(q_ptr), nq, float(radius), float(sigma),
                                                 1 if inclusive else 0, C.c_void_p(dens_ptr), C.c_void_p(cnt_ptr or 0),
                                                 C.c_void_p(stream or 0)))

    def rewire_candidates(self, space, x, k, resolution):
        """pomp_rewire_candidates_h: the k nearest nodes of x plus the feasibility of the straight edges
        node -> x (ok_in) and x -> node (ok_out), one call.  Returns (idx[c], dist[c], ok_in[c], ok_out[c])."""
        D = self.dim
        q = (C.c_double * D)(*[float(v) for v in x])
        idx = np.empty(k, dtype=np.int64)
        dist = np.empty(k, dtype=np.float64)
        oin = np.empty(k, dtype=np.uint8)
        oout = np.empty(k, dtype=np.uint8)
        cnt = C.c_int32()
        check(self.lib, self.lib.pomp_rewire_candidates_h(self._h, space._h, q, int(k), float(resolution),
                                                          idx.ctypes.data, dist.ctypes.data, C.byref(cnt),
                                                          oin.ctypes.data, oout.ctypes.data))
        c = cnt.value
        return idx[:c], dist[:c], oin[:c].astype(bool), oout[:c].astype(bool)

    def nearest_device(self, q_ptr, nq, idx_ptr, dist_ptr, stream=None):
        check(self.lib, self.lib.pomp_nn_nearest(self._h, C.c_void_p(q_ptr), nq, C.c_void_p(idx_ptr),
                                                 C.c_void_p(dist_ptr), C.c_void_p(stream or 0)))

    def knearest_device(self, q_ptr, nq, k, idx_ptr, dist_ptr, cnt_ptr=None, stream=None):
        check(self.lib, self.lib.pomp_nn_knearest(self._h, C.c_void_p(q_ptr), nq, k, C.c_void_p(idx_ptr),
                                                  C.c_void_p(dist_ptr), C.c_void_p(cnt_ptr or 0),
                                                  C.c_void_p(stream or 0)))

    def neighbors_device(self, q_ptr, nq, radius, inclusive, off_ptr, idx_ptr, idx_capacity, stream=None):
        need = C.c_int64()
        rc = self.lib.pomp_nn_neighbors(self._h, C.c_void_p(q_ptr), nq, float(radius), 1 if inclusive else 0,
                                        C.c_void_p(off_ptr), C.c_void_p(idx_ptr), idx_capacity, C.byref(need),
                                        C.c_void_p(stream or 0))
        if rc != _abi.POMP_ERR_CAPACITY:
            check(self.lib, rc)
        return rc, need.value


def merge_candidates_device(idx_in_ptr, dist_in_ptr, n_shards, nq, k, idx_out_ptr, dist_out_ptr, cnt_out_ptr=None,
                            stream=None):
    lib = _abi.load_library()
    check(lib, lib.pomp_nn_merge_candidates(C.c_void_p(idx_in_ptr), C.c_void_p(dist_in_ptr), n_shards, nq, k,
                                            C.c_void_p(idx_out_ptr), C.c_void_p(dist_out_ptr),
                                            C.c_void_p(cnt_out_ptr or 0), C.c_void_p(stream or 0)))


class CudaSpace(object):
    """Feasibility + edge checks on one GPU (pomp_space_*, pomp_feasible*, pomp_edge_check*)."""

    def __init__(self, space_spec, device=0):
        self.lib = _abi.load_library()
        self.spec = space_spec
        self.dim = space_spec.dim
        self.device = device
        self._h = C.c_void_p()
        d = space_spec.desc()
        check(self.lib, self.lib.pomp_space_create(C.byref(d), device, C.byref(self._h)))
        self._a1 = (C.c_double * _abi.MAX_DIM)()
        self._b1 = (C.c_double * _abi.MAX_DIM)()
        self._ok1 = C.c_uint8()

    def close(self):
        if self._h is not None and self._h.value:
            self.lib.pomp_space_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bounds(self, lo, hi):
        lo, hi = _f64(lo), _f64(hi)
        check(self.lib, self.lib.pomp_space_set_bounds(self._h, _ptr(lo), _ptr(hi)))

    def get_stat(self, name):
        out = C.c_double()
        check(self.lib, self.lib.pomp_space_get_stat(self._h, name.encode(), C.byref(out)))
        return out.value

    def feasible(self, x):
        x = _f64(x, self.dim)
        ok = np.empty(len(x), np.uint8)
        check(self.lib, self.lib.pomp_feasible_h(self._h, _ptr(x), len(x), _ptr(ok)))
        return ok.astype(bool)

    def feasible_one(self, x):
        a = self._a1
        for j in range(self.dim):
            a[j] = x[j]
        check(self.lib, self.lib.pomp_feasible_h(self._h, a, 1, C.byref(self._ok1)))
        return bool(self._ok1.value)

    def edge_check(self, a, b, resolution):
        a, b = _f64(a, self.dim), _f64(b, self.dim)
        ok = np.empty(len(a), np.uint8)
        check(self.lib, self.lib.pomp_edge_check_h(self._h, _ptr(a), _ptr(b), len(a), float(resolution), _ptr(ok)))
        return ok.astype(bool)

    def edge_check_one(self, a, b, resolution):
        pa, pb = self._a1, self._b1
        for j in range(self.dim):
            pa[j] = a[j]
            pb[j] = b[j]
        check(self.lib, self.lib.pomp_edge_check_h(self._h, pa, pb, 1, float(resolution), C.byref(self._ok1)))
        return bool(self._ok1.value)

    def edge_check_path(self, geodesic_spec, waypoints, offsets, resolution):
        wp = _f64(waypoints, self.dim)
        off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32))
        ok = np.empty(len(off) - 1, np.uint8)
        check(self.lib, self.lib.pomp_edge_check_path_h(self._h, C.byref(geodesic_spec.desc()), _ptr(wp), _ptr(off),
                                                        len(off) - 1, float(resolution), _ptr(ok)))
        return ok.astype(bool)

    def edge_check_control(self, dyn_spec, geodesic_spec, x, u, resolution):
        x, u = _f64(x, dyn_spec.x_dim), _f64(u, dyn_spec.u_dim)
        ok = np.empty(len(x), np.uint8)
        g = C.byref(geodesic_spec.desc()) if geodesic_spec is not None else None
        check(self.lib, self.lib.pomp_edge_check_control_h(self._h, C.byref(dyn_spec.desc()), g, _ptr(x), _ptr(u),
                                                           len(x), float(resolution), _ptr(ok)))
        return ok.astype(bool)

    # device-pointer variants
    def feasible_device(self, x_ptr, n, ok_ptr, stream=None):
        check(self.lib, self.lib.pomp_feasible(self._h, C.c_void_p(x_ptr), n, C.c_void_p(ok_ptr),
                                               C.c_void_p(stream or 0)))

    def edge_check_device(self, a_ptr, b_ptr, n, resolution, ok_ptr, stream=None):
        check(self.lib, self.lib.pomp_edge_check(self._h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), n, float(resolution),
                                                 C.c_void_p(ok_ptr), C.c_void_p(stream or 0)))

    def edge_check_indexed_device(self, nodes_ptr, ij_ptr, n_edges, resolution, ok_ptr, stream=None):
        check(self.lib, self.lib.pomp_edge_check_indexed(self._h, C.c_void_p(nodes_ptr), C.c_void_p(ij_ptr), n_edges,
                                                         float(resolution), C.c_void_p(ok_ptr),
                                                         C.c_void_p(stream or 0)))

    def edge_check_indexed_nodes_device(self, nodes_ptr, n_nodes, ij_ptr, n_edges, resolution, node_ok_ptr, ok_ptr,
                                        stream=None):
        """Roadmap edges with the node count known: feasibility once per node (flags returned in node_ok), the edges
        gather the flags and test their interior samples."""
        check(self.lib, self.lib.pomp_edge_check_indexed_nodes(self._h, C.c_void_p(nodes_ptr), n_nodes,
                                                               C.c_void_p(ij_ptr), n_edges, float(resolution),
                                                               C.c_void_p(node_ok_ptr), C.c_void_p(ok_ptr),
                                                               C.c_void_p(stream or 0)))


class CudaDynamics(object):
    """Batched nextState / fused control selection (pomp_next_state*, pomp_select_control*)."""

    def __init__(self, dyn_spec):
        self.lib = _abi.load_library()
        self.spec = dyn_spec
        self._desc = dyn_spec.desc()

    def next_state(self, x, u):
        X = self.spec.state_dim()
        x, u = _f64(x, X), _f64(u, self.spec.u_dim)
        out = np.empty_like(x)
        check(self.lib, self.lib.pomp_next_state_h(C.byref(self._desc), _ptr(x), _ptr(u), len(x), _ptr(out)))
        return out

    def select_control(self, metric_spec, x, xdes, u, valid=None):
        X = self.spec.state_dim()
        x, xdes = _f64(x, X), _f64(xdes, X)
        B = len(x)
        u = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(B, -1, self.spec.u_dim))
        S = u.shape[1]
        v = None if valid is None else np.ascontiguousarray(np.asarray(valid, dtype=np.uint8).reshape(B, S))
        best = np.empty(B, np.int32)
        xb = np.empty((B, X), np.float64)
        db = np.empty(B, np.float64)
        check(self.lib, self.lib.pomp_select_control_h(C.byref(self._desc), C.byref(metric_spec.desc()), _ptr(x),
                                                       _ptr(xdes), _ptr(u), _ptr(v) if v is not None else None, B, S,
                                                       _ptr(best), _ptr(xb), _ptr(db)))
        return best, xb, db

    def next_state_device(self, x_ptr, u_ptr, n, out_ptr, stream=None):
        check(self.lib, self.lib.pomp_next_state(C.byref(self._desc), C.c_void_p(x_ptr), C.c_void_p(u_ptr), n,
                                                 C.c_void_p(out_ptr), C.c_void_p(stream or 0)))

    def select_control_device(self, metric_spec, x_ptr, xdes_ptr, u_ptr, valid_ptr, B, S, best_ptr, xbest_ptr,
                              dbest_ptr, stream=None):
        check(self.lib, self.lib.pomp_select_control(C.byref(self._desc), C.byref(metric_spec.desc()),
                                                     C.c_void_p(x_ptr), C.c_void_p(xdes_ptr), C.c_void_p(u_ptr),
                                                     C.c_void_p(valid_ptr or 0), B, S, C.c_void_p(best_ptr),
                                                     C.c_void_p(xbest_ptr), C.c_void_p(dbest_ptr),
                                                     C.c_void_p(stream or 0)))


def rrt_extend(nn, space, xrand, max_step, resolution, check_sample=True, _buf={}):
    """pomp_rrt_extend_h: returns (parent_index, xnew list, edge_length, added)."""
    lib = nn.lib
    D = nn.dim
    if D not in _buf:
        _buf[D] = ((C.c_double * D)(), (C.c_double * D)(), C.c_int64(), C.c_int32(), C.c_double())
    q, xn, parent, added, elen = _buf[D]
    for j in range(D):
        q[j] = xrand[j]
    check(lib, lib.pomp_rrt_extend_h(nn._h, space._h, q, 1 if check_sample else 0, float(max_step),
                                     float(resolution), C.byref(parent), xn, C.byref(elen), C.byref(added)))
    return parent.value, list(xn), elen.value, bool(added.value)


class CudaForest(object):
    """T independent kinematic RRTs advanced in lock step, one launch per iteration
    (pomp_forest_* entry points)."""

    def __init__(self, dim, n_trials, capacity=16384, device=0):
        self.lib = _abi.load_library()
        self.dim, self.T = dim, n_trials
        self._h = C.c_void_p()
        check(self.lib, self.lib.pomp_forest_create(dim, n_trials, capacity, device, C.byref(self._h)))
        T, D = n_trials, dim
        self.xrand = (C.c_double * (T * D))()
        self.root = (C.c_double * D)()
        self.active = (C.c_uint8 * T)()
        self.reset = (C.c_uint8 * T)()
        self.parent = (C.c_int64 * T)()
        self.xnew = (C.c_double * (T * D))()
        self.elen = (C.c_double * T)()
        self.added = (C.c_int32 * T)()

    def close(self):
        if self._h is not None and self._h.value:
            self.lib.pomp_forest_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self, trial):
        return self.lib.pomp_forest_size(self._h, trial)

    def extend(self, space, max_step, resolution, check_sample=True):
        """Inputs are read from self.xrand / self.root / self.active / self.reset, results land in
        self.parent / self.xnew / self.elen / self.added."""
        check(self.lib, self.lib.pomp_forest_extend_h(self._h, space._h, self.xrand, self.root, self.active,
                                                      self.reset, 1 if check_sample else 0, float(max_step),
                                                      float(resolution), self.parent, self.xnew, self.elen,
                                                      self.added))


    MAX_MULTI = 32

    def extend_multi(self, space, n_iters, max_step, resolution, check_sample=True, goal=None, goal_radius=0.0):
        """pomp_forest_extend_multi_h: n_iters expansions per trial in one launch.  Samples are read from
        self.xrand_m ([T][n_iters][D], row-major), results land in self.n_done / parent_m / xnew_m / elen_m /
        added_m (only the first n_done[t] records of trial t are valid)."""
        T, D, M = self.T, self.dim, int(n_iters)
        if getattr(self, "_m_alloc", 0) < M:
            self.xrand_m = (C.c_double * (T * self.MAX_MULTI * D))()
            self.n_done = (C.c_int32 * T)()
            self.parent_m = (C.c_int64 * (T * self.MAX_MULTI))()
            self.xnew_m = (C.c_double * (T * self.MAX_MULTI * D))()
            self.elen_m = (C.c_double * (T * self.MAX_MULTI))()
            self.added_m = (C.c_int32 * (T * self.MAX_MULTI))()
            self._m_alloc = self.MAX_MULTI
        g = None
        if goal is not None:
            g = (C.c_double * D)(*[float(v) for v in goal])
        check(self.lib, self.lib.pomp_forest_extend_multi_h(self._h, space._h, M, self.xrand_m, self.root, self.active,
                                                            self.reset, 1 if check_sample else 0, float(max_step),
                                                            float(resolution), g, float(goal_radius), self.n_done,
                                                            self.parent_m, self.xnew_m, self.elen_m, self.added_m))


class CudaProjectionDensity(object):
    """ESTWithProjections' hash density on the GPU (pomp_est_* entry points, csrc/est.cu).
    bases: list of (indices[<=4], scales[<=4], weight)."""

    def __init__(self, dim, bases, device=0):
        self.lib = _abi.load_library()
        self.dim = dim
        nb = len(bases)
        bd = np.array([len(b[0]) for b in bases], np.int32)
        bi = np.zeros((nb, 4), np.int32)
        bs = np.zeros((nb, 4), np.float64)
        bw = np.array([float(b[2]) for b in bases], np.float64)
        for j, (idx, sc, _) in enumerate(bases):
            bi[j, :len(idx)] = idx
            bs[j, :len(sc)] = sc
        self._h = C.c_void_p()
        check(self.lib, self.lib.pomp_est_create(dim, nb, _ptr(bd), _ptr(bi), _ptr(bs), _ptr(bw), device, C.byref(self._h)))

    def close(self):
        if self._h is not None and self._h.value:
            self.lib.pomp_est_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(self.lib, self.lib.pomp_est_reset(self._h))

    def size(self):
        return self.lib.pomp_est_size(self._h)

    def add(self, pts):
        pts = _f64(pts, self.dim)
        check(self.lib, self.lib.pomp_est_add_h(self._h, _ptr(pts), len(pts)))

    def density(self, q):
        q = _f64(q, self.dim)
        out = np.empty(len(q), np.float64)
        check(self.lib, self.lib.pomp_est_density_h(self._h, _ptr(q), len(q), _ptr(out)))
        return out


class CudaSSTForest(object):
    """T lock-step sst* trials on one GPU (pomp_sst_* entry points, csrc/sst.cu); the host driver is sst.SSTStarForest."""

    def __init__(self, x_dim, u_dim, n_trials, node_cap, witness_cap, dyn_spec, geodesic_spec, space_spec, device=0):
        self.lib = _abi.load_library()
        self.X, self.U, self.T, self.ncap = x_dim, u_dim, n_trials, node_cap
        self.space = CudaSpace(space_spec, device)
        self._h = C.c_void_p()
        self._dyn, self._geo = dyn_spec.desc(), geodesic_spec.desc()
        check(self.lib, self.lib.pomp_sst_create(x_dim, u_dim, n_trials, node_cap, witness_cap, C.byref(self._dyn),
                                                 C.byref(self._geo), device, C.byref(self._h)))
        self._best = np.empty(n_trials, np.float64)
        self._nodes = np.empty(n_trials, np.int32)
        self._none = np.empty(n_trials, np.int32)

    def close(self):
        if self._h is not None and self._h.value:
            self.lib.pomp_sst_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, root):
        r = _f64(root)
        check(self.lib, self.lib.pomp_sst_reset_h(self._h, _ptr(r)))

    def step(self, M, xr, uu, valid, sel, wit, restart, goal_c, goal_r, res, restart_now=False, half_sq=None):
        """half_sq = (per-sample 0.5*dt_last**2 [T][M], 0.5*dt**2) computed by the caller with Python's ** (CV dynamics)."""
        xr = _f64(xr)
        uu = _f64(uu)
        valid = np.ascontiguousarray(np.asarray(valid, dtype=np.uint8))
        sel, wit = _f64(sel if M else [0.0]), _f64(wit if M else [0.0])
        rs = np.ascontiguousarray(np.asarray(restart if M else [0], dtype=np.uint8))
        gc = _f64(goal_c)
        check(self.lib, self.lib.pomp_sst_step_h(self._h, self.space._h, int(M), _ptr(xr), _ptr(uu), _ptr(valid), _ptr(sel),
                                                 _ptr(wit), _ptr(rs), 1 if restart_now else 0,
                                                 _ptr(gc) if goal_r >= 0 else None, float(goal_r), float(res),
                                                 _ptr(_f64(half_sq[0])) if (half_sq is not None and M) else None,
                                                 float(half_sq[1]) if half_sq is not None else 0.0,
                                                 _ptr(self._best), _ptr(self._nodes), _ptr(self._none)))
        return self._best.copy(), self._nodes.copy(), self._none.copy()

    def get_trial(self, trial):
        cnt = C.c_int32()
        x = np.empty((self.ncap, self.X), np.float64)
        parent = np.empty(self.ncap, np.int32)
        alive = np.empty(self.ncap, np.uint8)
        check(self.lib, self.lib.pomp_sst_get_trial_h(self._h, int(trial), _ptr(x), _ptr(parent), _ptr(alive), None, None,
                                                      C.byref(cnt)))
        n = cnt.value
        return x[:n].tolist(), parent[:n].tolist(), [bool(a) for a in alive[:n]]


def launch_count():
    return _abi.load_library().pomp_launch_count()


def device_count():
    return _abi.load_library().pomp_device_count()
