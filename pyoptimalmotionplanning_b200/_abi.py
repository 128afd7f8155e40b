"""ctypes mirror of include/pomp_b200.h and the loader of the C-ABI library.

The library (csrc/libpomp_b200.so) is the product: hand-written sm_100a CUDA kernels behind
extern "C" entry points.  There is no CPU fallback -- if the library is missing, or no CUDA
device is visible, the product raises instead of computing something else.
"""
import ctypes as C
import os

MAX_DIM = 8
MAX_COMP = 8

# pomp_metric_kind
METRIC_EUCLIDEAN, METRIC_WEIGHTED_EUCLIDEAN, METRIC_SCALED_EUCLIDEAN, METRIC_L1, METRIC_LINF, \
    METRIC_GEODESIC_PRODUCT = range(6)
COMP_CARTESIAN, COMP_SO2 = 0, 1
# pomp_dyn_kind
DYN_ADAPTOR, DYN_CV, DYN_CV_EULER, DYN_DUBINS, DYN_PENDULUM, DYN_LTI, DYN_FLAPPY = range(7)
# pomp_cost_kind
COST_NONE, COST_TIME, COST_ABS_U0, COST_PATH_LENGTH = range(4)

POMP_OK, POMP_ERR_INVALID, POMP_ERR_CUDA, POMP_ERR_CAPACITY, POMP_ERR_NOMEM = 0, -1, -2, -3, -4


class MetricDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dim", C.c_int32), ("n_comp", C.c_int32),
        ("comp_kind", C.c_int32 * MAX_COMP), ("comp_dim", C.c_int32 * MAX_COMP),
        ("comp_weight", C.c_double * MAX_COMP), ("dim_weight", C.c_double * MAX_DIM),
        ("scale", C.c_double),
        ("has_cost", C.c_int32), ("cost_weight_set", C.c_int32), ("cost_max_set", C.c_int32),
        ("_pad", C.c_int32),
        ("cost_weight", C.c_double), ("cost_max", C.c_double),
    ]


class SpaceDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("obstacle_dim0", C.c_int32), ("obstacle_dim1", C.c_int32),
        ("n_boxes", C.c_int32), ("n_circles", C.c_int32), ("_pad", C.c_int32),
        ("lo", C.c_double * MAX_DIM), ("hi", C.c_double * MAX_DIM),
        ("boxes_h", C.c_void_p), ("circles_h", C.c_void_p),
    ]


class DynDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("x_dim", C.c_int32), ("u_dim", C.c_int32), ("cost_kind", C.c_int32),
        ("dt", C.c_double), ("p", C.c_double * 8),
        ("A", C.c_double * (MAX_DIM * MAX_DIM)), ("B", C.c_double * (MAX_DIM * MAX_DIM)),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_D = C.c_double
_MD = C.POINTER(MetricDesc)
_SD = C.POINTER(SpaceDesc)
_DD = C.POINTER(DynDesc)

# name -> (restype, argtypes); every symbol include/pomp_b200.h declares
SIGNATURES = {
    "pomp_last_error": (C.c_char_p, []),
    "pomp_version": (C.c_int, []),
    "pomp_device_count": (C.c_int, []),
    "pomp_launch_count": (_I64, []),
    "pomp_fp64_unfused_peak": (C.c_int, [C.c_int, _P]),
    "pomp_nn_create": (C.c_int, [_MD, _I64, C.c_int, C.POINTER(_P)]),
    "pomp_nn_destroy": (C.c_int, [_P]),
    "pomp_nn_reset": (C.c_int, [_P]),
    "pomp_nn_add": (C.c_int, [_P, _P, _I64, _P, _P]),
    "pomp_nn_add_h": (C.c_int, [_P, _P, _I64, _P]),
    "pomp_nn_set": (C.c_int, [_P, _P, _I64, _P]),
    "pomp_nn_set_h": (C.c_int, [_P, _P, _I64]),
    "pomp_nn_remove": (C.c_int, [_P, _I64]),
    "pomp_nn_set_mask_h": (C.c_int, [_P, _P, _I64]),
    "pomp_nn_update_metric": (C.c_int, [_P, _MD]),
    "pomp_nn_size": (_I64, [_P]),
    "pomp_nn_set_param": (C.c_int, [_P, C.c_char_p, _D]),
    "pomp_nn_get_stat": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double)]),
    "pomp_nn_build_index": (C.c_int, [_P, _P]),
    "pomp_nn_nearest": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "pomp_nn_nearest_h": (C.c_int, [_P, _P, _I64, _P, _P]),
    "pomp_nn_knearest": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _P]),
    "pomp_nn_knearest_h": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P]),
    "pomp_nn_neighbors": (C.c_int, [_P, _P, _I64, _D, C.c_int, _P, _P, _I64, _P, _P]),
    "pomp_nn_neighbors_h": (C.c_int, [_P, _P, _I64, _D, C.c_int, _P, _P, _I64, _P]),
    "pomp_nn_merge_candidates": (C.c_int, [_P, _P, _I32, _I64, _I32, _P, _P, _P, _P]),
    "pomp_shard_route": (C.c_int, [_P, _I64, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "pomp_shard_finalize": (C.c_int, [_P, _I64, _I32, _I32, _D, _D, _I32, _P, _P, _P, _P, _I32, _P, _P, _P]),
    "pomp_shard_unpack": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P, _P]),
    "pomp_shard_route_fixed": (C.c_int, [_P, _I64, _I32, _I32, _P, _I32, _I64, _P, _P, _P, _P, _I64, _P]),
    "pomp_shard_finalize_rows": (C.c_int, [_P, _I64, _I32, _I32, _D, _D, _I32, _P, _P, _P, _P, _I32, _P, _P]),
    "pomp_shard_unpack_rows": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P, _P, _I64, _P]),
    "pomp_shard_route_p2p": (C.c_int, [_P, _I64, _I32, _I32, _P, _I32, _I32, _I64, _P, _P, _P, _P, _I64, _P]),
    "pomp_shard_finalize_p2p": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _D, _D, _I32, _P, _P, _P, _P, _P, _P]),
    "pomp_shard_share_counts_p2p": (C.c_int, [_P, _I32, _P, _I64, _I32, _P]),
    "pomp_shard_worst_count": (C.c_int, [_P, _I32, _I32, _P, _P]),
    "pomp_shard_unpack_p2p": (C.c_int, [_P, _I32, _I64, _I32, _P, _P, _P, _P, _P, _I64, _P]),
    "pomp_shard_gather_deferred": (C.c_int, [_P, _I32, _P, _P, _I64, _P, _P, _P]),
    "pomp_shard_scatter_deferred": (C.c_int, [_P, _P, _I32, _P, _P, _I64, _P, _P, _P]),
    "pomp_edge_check_indexed_nodes": (C.c_int, [_P, _P, _I64, _P, _I64, C.c_double, _P, _P, _P]),
    "pomp_roadmap_build": (C.c_int, [_P, _P, _I32, C.c_double, _I32, _P, _P, _I64, _P, _P, _P]),
    "pomp_roadmap_build_h": (C.c_int, [_P, _P, _I32, C.c_double, _I32, _P, _P, _I64, _P, _P]),
    "pomp_est_create": (C.c_int, [_I32, _I32, _P, _P, _P, _P, C.c_int, C.POINTER(_P)]),
    "pomp_est_destroy": (C.c_int, [_P]),
    "pomp_est_reset": (C.c_int, [_P]),
    "pomp_est_size": (_I64, [_P]),
    "pomp_est_add_h": (C.c_int, [_P, _P, _I64]),
    "pomp_est_density": (C.c_int, [_P, _P, _I64, _P, _P]),
    "pomp_est_density_h": (C.c_int, [_P, _P, _I64, _P]),
    "pomp_sst_create": (C.c_int, [_I32, _I32, _I32, _I32, _I32, _P, _P, C.c_int, C.POINTER(_P)]),
    "pomp_sst_destroy": (C.c_int, [_P]),
    "pomp_sst_reset_h": (C.c_int, [_P, _P]),
    "pomp_sst_step_h": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _P, _P, _I32, _P, _D, _D, _P, _D, _P, _P, _P]),
    "pomp_sst_get_trial_h": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P]),
    "pomp_space_create": (C.c_int, [_SD, C.c_int, C.POINTER(_P)]),
    "pomp_space_destroy": (C.c_int, [_P]),
    "pomp_space_set_bounds": (C.c_int, [_P, _P, _P]),
    "pomp_space_get_stat": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double)]),
    "pomp_feasible": (C.c_int, [_P, _P, _I64, _P, _P]),
    "pomp_feasible_h": (C.c_int, [_P, _P, _I64, _P]),
    "pomp_edge_check": (C.c_int, [_P, _P, _P, _I64, _D, _P, _P]),
    "pomp_edge_check_h": (C.c_int, [_P, _P, _P, _I64, _D, _P]),
    "pomp_edge_check_indexed": (C.c_int, [_P, _P, _P, _I64, _D, _P, _P]),
    "pomp_edge_check_path": (C.c_int, [_P, _MD, _P, _P, _I64, _D, _P, _P]),
    "pomp_edge_check_path_h": (C.c_int, [_P, _MD, _P, _P, _I64, _D, _P]),
    "pomp_edge_check_control": (C.c_int, [_P, _DD, _MD, _P, _P, _I64, _D, _P, _P]),
    "pomp_edge_check_control_h": (C.c_int, [_P, _DD, _MD, _P, _P, _I64, _D, _P]),
    "pomp_next_state": (C.c_int, [_DD, _P, _P, _I64, _P, _P]),
    "pomp_next_state_h": (C.c_int, [_DD, _P, _P, _I64, _P]),
    "pomp_select_control": (C.c_int, [_DD, _MD, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P]),
    "pomp_select_control_h": (C.c_int, [_DD, _MD, _P, _P, _P, _P, _I64, _I32, _P, _P, _P]),
    "pomp_forest_create": (C.c_int, [C.c_int, C.c_int, _I64, C.c_int, C.POINTER(_P)]),
    "pomp_forest_destroy": (C.c_int, [_P]),
    "pomp_forest_size": (_I64, [_P, C.c_int]),
    "pomp_forest_extend_h": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, _D, _D, _P, _P, _P, _P]),
    "pomp_forest_extend_multi_h": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, C.c_int, _D, _D, _P, _D, _P, _P, _P, _P,
                                             _P]),
    "pomp_rrt_extend_h": (C.c_int, [_P, _P, _P, C.c_int, _D, _D, _P, _P, _P, _P]),
    "pomp_nn_density": (C.c_int, [_P, _P, C.c_int64, _D, _D, C.c_int, _P, _P, _P]),
    "pomp_nn_density_h": (C.c_int, [_P, _P, C.c_int64, _D, _D, C.c_int, _P, _P]),
    "pomp_rewire_candidates_h": (C.c_int, [_P, _P, _P, C.c_int32, _D, _P, _P, _P, _P, _P]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libpomp_b200.so")


class PompError(RuntimeError):
    """A C-ABI call returned a non-zero pomp_status."""

    def __init__(self, status, message):
        RuntimeError.__init__(self, "pomp_b200 status %d: %s" % (status, message))
        self.status = status


_lib = None


def load_library(path=None):
    """dlopen the C-ABI library and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            "pomp_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(lib, status):
    if status != 0:
        msg = lib.pomp_last_error()
        raise PompError(status, msg.decode() if msg else "")
    return status
