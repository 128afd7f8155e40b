/*
 * pomp_b200.h -- C-ABI of the B200-native hot path of pyOptimalMotionPlanning (pomp).
 *
 * The reference (krishauser/pyOptimalMotionPlanning) is pure Python and has no FFI; its
 * "plugin boundary" is duck typing on four object families.  Every entry point below names
 * the reference interface it replaces (file:line relative to the reference checkout):
 *
 *   pomp_nn_*        <- pomp/structures/nearestneighbors.py:13-141  (NearestNeighbors)
 *   pomp_feasible    <- pomp/spaces/configurationspace.py:142,196 ; example_problems/geometric.py:56-60
 *   pomp_edge_check* <- pomp/spaces/edgechecker.py:20-30            (EpsilonEdgeChecker.feasible)
 *   pomp_next_state  <- pomp/spaces/controlspace.py:78,249-267 ; statespace.py:62-70 ; dubins.py:107-125
 *   pomp_select_control <- pomp/planners/kinodynamicplanner.py:66-84 (RandomControlSelector.select)
 *   pomp_rrt_extend  <- pomp/planners/kinodynamicplanner.py:354-411 (RRT.expand, kinematic case)
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy / C++ types cross this boundary.
 *   - every function returns 0 on success, a negative pomp_status otherwise;
 *     pomp_last_error() returns the message of the calling thread's last failure.
 *   - pointers without a suffix are DEVICE pointers (any CUDA allocation of the current
 *     device, e.g. a torch tensor's data_ptr()); pointers suffixed _h are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device
 *     entry points only enqueue work; the *_h variants synchronise before returning.
 *   - all arithmetic is IEEE float64, one rounding per operation, no FMA contraction, in the
 *     reference's operation order.  Indices are insertion indices.
 *   - ties are broken by lowest insertion index; k-NN results are ordered by (distance, index),
 *     radius results by index.
 *   - handles are not thread-safe; the caller owns every buffer it passes.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     POMP_ERR_CUDA.
 */
#ifndef POMP_B200_H
#define POMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POMP_MAX_DIM 8   /* state dimension incl. the cost coordinate (reference configs use 2..7) */
#define POMP_MAX_COMP 8  /* components of a product space */

typedef enum {
  POMP_OK = 0,
  POMP_ERR_INVALID = -1,  /* bad argument / unsupported descriptor */
  POMP_ERR_CUDA = -2,     /* CUDA runtime failure (incl. no device)  */
  POMP_ERR_CAPACITY = -3, /* caller-provided output too small        */
  POMP_ERR_NOMEM = -4
} pomp_status;

/* ---------------------------------------------------------------- metrics -------------- */
/* pomp/spaces/metric.py:4-23, pomp/spaces/geodesicspace.py:45-52, kinodynamicplanner.py:899-913 */
typedef enum {
  POMP_METRIC_EUCLIDEAN = 0,          /* vectorops.distance: sqrt(sum (a-b)*(a-b))                    */
  POMP_METRIC_WEIGHTED_EUCLIDEAN = 1, /* WeightedEuclideanMetric(w list): sqrt(sum w_i*(a_i-b_i)**2)   */
  POMP_METRIC_SCALED_EUCLIDEAN = 2,   /* WeightedEuclideanMetric(w scalar): w*euclidean               */
  POMP_METRIC_L1 = 3,
  POMP_METRIC_LINF = 4,
  POMP_METRIC_GEODESIC_PRODUCT = 5    /* MultiGeodesicSpace.distance: sqrt(sum_c dist_c**2 * w_c)     */
} pomp_metric_kind;

typedef enum { POMP_COMP_CARTESIAN = 0, POMP_COMP_SO2 = 1 } pomp_comp_kind;

typedef struct {
  int32_t kind;                      /* pomp_metric_kind                                             */
  int32_t dim;                       /* full point dimension, INCLUDING the cost coordinate if has_cost */
  int32_t n_comp;                    /* GEODESIC_PRODUCT: number of components                       */
  int32_t comp_kind[POMP_MAX_COMP];  /* pomp_comp_kind                                               */
  int32_t comp_dim[POMP_MAX_COMP];   /* coordinates per component (SO2: 1)                           */
  double comp_weight[POMP_MAX_COMP]; /* MultiGeodesicSpace.componentWeights                          */
  double dim_weight[POMP_MAX_DIM];   /* WEIGHTED_EUCLIDEAN                                           */
  double scale;                      /* SCALED_EUCLIDEAN                                             */
  /* CostMetric wrapper (kinodynamicplanner.py:899-913): last coordinate is accumulated cost;
     a = tree node, b = query.  d = base(a[:-1],b[:-1]); if cost_weight_set: inf when cost_max_set
     and (a_c > cost_max or b_c > cost_max); d += max(a_c-b_c,0)*cost_weight.                       */
  int32_t has_cost;
  int32_t cost_weight_set;
  int32_t cost_max_set;
  int32_t _pad;
  double cost_weight;
  double cost_max;
} pomp_metric_desc;

/* ---------------------------------------------------------------- nearest neighbours ---- */
typedef struct pomp_nn pomp_nn;

const char* pomp_last_error(void);
int pomp_version(void);
/* number of CUDA devices visible (0 when there is none; never fails) */
int pomp_device_count(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches) */
int64_t pomp_launch_count(void);
/* diagnostics: sustained FP64 operations/s of separate DADD/DMUL (no FMA) on `device` -- the roofline of the
   brute-force kernels, which must round after every operation like the reference (pomp/klampt/vectorops.py:94-102) */
int pomp_fp64_unfused_peak(int device, double* flops_out);

/* NearestNeighbors.__init__ (nearestneighbors.py:14-23) */
int pomp_nn_create(const pomp_metric_desc* metric, int64_t capacity_hint, int device, pomp_nn** out);
int pomp_nn_destroy(pomp_nn* h);
/* NearestNeighbors.reset (:25-30) */
int pomp_nn_reset(pomp_nn* h);
/* NearestNeighbors.add (:32-39): append n points (row-major n x dim); indices are insertion order */
int pomp_nn_add(pomp_nn* h, const double* pts, int64_t n, int64_t* first_index_out_h, void* stream);
int pomp_nn_add_h(pomp_nn* h, const double* pts_h, int64_t n, int64_t* first_index_out_h);
/* NearestNeighbors.set (:65-74): bulk replace + index rebuild */
int pomp_nn_set(pomp_nn* h, const double* pts, int64_t n, void* stream);
int pomp_nn_set_h(pomp_nn* h, const double* pts_h, int64_t n);
/* NearestNeighbors.remove (:41-62): tombstone by insertion index (the host binding resolves
   (point,data) -> index, exactly-equal-coordinates semantics stay on the host) */
int pomp_nn_remove(pomp_nn* h, int64_t index);
/* filter-as-mask: reject_mask_h[i] != 0 means the reference's filter(point,data) returned True.
   NULL clears the mask.  n must equal pomp_nn_size(). */
int pomp_nn_set_mask_h(pomp_nn* h, const uint8_t* reject_mask_h, int64_t n);
/* CostSpaceRRT.updateBestCost mutates the metric in place (kinodynamicplanner.py:1059-1077) */
int pomp_nn_update_metric(pomp_nn* h, const pomp_metric_desc* metric);
int64_t pomp_nn_size(const pomp_nn* h);
/* Tuning knobs (every name pomp_nn_set_param accepts; defaults in nn_handle.cuh; none of them changes an answer):
     regime        "grid_min_points" (16384), "grid_min_queries" (64): below either, the brute-force kernels;
                   "max_tail" (2048): points appended since the last index build before it is rebuilt;
                   "points_per_cell" (0 = 1.5 / 2 / 4 by dimension): target cell occupancy, rebuilds the index;
                   "periodic_axes" (1): SO(2) components of a product metric become wrap-around grid axes
     query sort    "sort_queries" (1), "sort_min_queries" (32768), "sort_shift": bucket sort of large batches by
                   coarse home bin; "overlap_chunks" (1), "overlap_min_queries": sort of chunk c+1 on a second
                   stream (measured slower, kept off)
     kernel choice "strip_search" (1), "strip_min_per_tile", "strip_spin_ns", "strip_probe": streaming kernel for dense
                   2-D nearest() batches; "block_search" (1; 2 = also in 3-D), "block_need": fixed block of cells +
                   certificate in 2-D; "warp_search" (1 = where it wins, 2 = always, 0 = never), "warp_lanes" (32; 8,
                   16), "warp_need": the warp-per-query shell kernel; "se2_fast" (1): compile-time SE(2) kernel;
                   "bf_tiled" (1): shared-memory tiled brute force for Euclidean k = 1 batches; "single_radius" (1):
                   one-launch path of a single neighbors() query in the brute-force regime
     host pipeline "pipeline_chunks" (6), "pipeline_shape" (1 = triangle), "use_graphs" (1): the *_h entry points
     measurement   "profile" (0): CUDA events around the dominant search kernel, read back with
                   pomp_nn_get_stat("search_kernel_ms" / "search_kernel_launches") */
int pomp_nn_set_param(pomp_nn* h, const char* name, double value);
/* read-outs: "search_kernel_ms" / "search_kernel_launches" (CUDA-event time of the dominant search
   kernel since set_param("profile",1); call after synchronising), "n_cells", "n_indexed", "grid_axes" */
int pomp_nn_get_stat(pomp_nn* h, const char* name, double* out);
/* (re)build the uniform-grid bucket index over all points now (otherwise built lazily) */
int pomp_nn_build_index(pomp_nn* h, void* stream);

/* NearestNeighbors.nearest (:76-96): idx = -1 and dist = +inf when no admissible point exists.  The reference returns
   the node only, so dist_h of the *_h variants may be NULL (the distances then stay on the device). */
int pomp_nn_nearest(pomp_nn* h, const double* q, int64_t nq, int64_t* idx, double* dist, void* stream);
int pomp_nn_nearest_h(pomp_nn* h, const double* q_h, int64_t nq, int64_t* idx_h, double* dist_h);
/* NearestNeighbors.knearest (:98-116) + KNearestResult (knn.py:7-30): per query the k smallest
   (dist, idx) ascending; unused slots are idx=-1, dist=+inf; count[q] = number found */
int pomp_nn_knearest(pomp_nn* h, const double* q, int64_t nq, int32_t k, int64_t* idx, double* dist,
                     int32_t* count, void* stream);
int pomp_nn_knearest_h(pomp_nn* h, const double* q_h, int64_t nq, int32_t k, int64_t* idx_h,
                       double* dist_h, int32_t* count_h);
/* NearestNeighbors.neighbors (:118-141).  inclusive=1: d <= radius (kd-tree branch, kdtree.py:347),
   inclusive=0: d < radius (brute-force branch :139).  CSR output: offsets[nq+1], idx sorted ascending
   per query.  If the hits do not fit idx_capacity the call fills only offsets and *needed_h and
   returns POMP_ERR_CAPACITY. */
int pomp_nn_neighbors(pomp_nn* h, const double* q, int64_t nq, double radius, int inclusive,
                      int64_t* offsets, int64_t* idx, int64_t idx_capacity, int64_t* needed_h,
                      void* stream);
int pomp_nn_neighbors_h(pomp_nn* h, const double* q_h, int64_t nq, double radius, int inclusive,
                        int64_t* offsets_h, int64_t* idx_h, int64_t idx_capacity, int64_t* needed_h);

/* multi-GPU, SPATIAL sharding (no reference counterpart: pomp is single-process).  Rank r owns the slab
   bounds[r-1] <= x[axis] < bounds[r] of the node set (bounds[-1] = -inf, bounds[n_ranks-1] = +inf) plus a halo copy of
   its neighbours' points near the slab faces.  pomp_shard_route counting-sorts a query batch by owning rank
   (q_out: queries grouped by rank, perm_out[t] = original position, counts_h[r] = queries for rank r; synchronises
   the stream: the host needs the split sizes for the all-to-all).  dest_scratch: nq int32, counts_dev: 128 uint64.
   pomp_shard_finalize turns local k-NN answers into packed 16-byte records (distance bits, GLOBAL index) for the
   return exchange and flags the queries whose k-th distance reaches outside the covered region [lo_cov, hi_cov]
   (uncertified[i] = 1); drop_halo = 1 blanks halo copies (the all-ranks fallback merges owned points only).
   pomp_shard_unpack writes records back to (idx, dist, count) at perm[i] (perm may be NULL). */
int pomp_shard_route(const double* q, int64_t nq, int32_t dim, int32_t axis, const double* bounds_h, int32_t n_ranks,
                     int32_t* dest_scratch, uint64_t* counts_dev, int64_t* counts_h, double* q_out, int64_t* perm_out,
                     void* stream);
int pomp_shard_finalize(const double* q, int64_t nq, int32_t dim, int32_t axis, double lo_cov, double hi_cov,
                        int32_t k, const int64_t* idx_local, const double* dist, const int64_t* local_to_global,
                        const uint8_t* halo_flag, int32_t drop_halo, int64_t* records, uint8_t* uncertified,
                        void* stream);
int pomp_shard_unpack(const int64_t* records, int64_t nq, int32_t k, const int64_t* perm, int64_t* idx, double* dist,
                      int32_t* count, void* stream);
/* The same exchange WITHOUT host synchronisation (equal-split all-to-alls): every destination rank gets a fixed block
   of `cap` query slots (q_out [n_ranks][cap][dim] / perm_out [n_ranks][cap], pre-filled by the caller with an ordinary
   point of the destination slab / -1: padding must be cheap to answer, a NaN query scans a whole shard);
   what does not fit is appended to defer_list (positions in the caller's batch; counters[64] counts them,
   counters[0..63] are the per-rank cursors: the caller zeroes all 65).  pomp_shard_finalize_rows writes rows of
   2k+1 words (k records + "no certificate" flag); pomp_shard_unpack_rows stores certified rows at perm[i] and appends
   the flagged ones to the deferred list; gather / scatter move the first F deferred queries into / out of the
   fixed-size block of the all-ranks fallback (padded with pad_point[dim], a device pointer). */
int pomp_shard_route_fixed(const double* q, int64_t nq, int32_t dim, int32_t axis, const double* bounds_h,
                           int32_t n_ranks, int64_t cap, uint64_t* counters, double* q_out, int64_t* perm_out,
                           int64_t* defer_list, int64_t defer_cap, void* stream);
int pomp_shard_finalize_rows(const double* q, int64_t nq, int32_t dim, int32_t axis, double lo_cov, double hi_cov,
                             int32_t k, const int64_t* idx_local, const double* dist, const int64_t* local_to_global,
                             const uint8_t* halo_flag, int32_t drop_halo, int64_t* rows, void* stream);
int pomp_shard_unpack_rows(const int64_t* rows, int64_t n, int32_t k, const int64_t* perm, int64_t* idx, double* dist,
                           int64_t* defer_list, uint64_t* defer_count, int64_t defer_cap, void* stream);
int pomp_shard_gather_deferred(const double* q, int32_t dim, const int64_t* defer_list, const uint64_t* defer_count,
                               int64_t F, const double* pad_point, double* qf, void* stream);
/* Peer-memory variant: the exchange is done by the kernels themselves with NVLink stores into buffers every rank
   maps from its peers (peer_*_blocks_h[r] = device address of rank r's block in THIS process, from CUDA IPC /
   torch symmetric memory).  route: query i owned by rank r is stored at rank r's block, row my_rank*cap + slot;
   finalize: the answer row of slot (src, t) is stored at rank src's result block, row my_rank*cap + t; unpack reads
   the local result block [n_ranks][cap] (slots below the per-rank cursors in counters[0..n_ranks) are valid).  The
   caller separates the three with device barriers; no collective is on the data path. */
int pomp_shard_route_p2p(const double* q, int64_t nq, int32_t dim, int32_t axis, const double* bounds_h, int32_t n_ranks,
                         int32_t my_rank, int64_t cap, uint64_t* counters, const uint64_t* peer_query_blocks_h,
                         int64_t* perm_out, int64_t* defer_list, int64_t defer_cap, void* stream);
int pomp_shard_finalize_p2p(const double* q, int64_t cap, int32_t n_ranks, int32_t my_rank, int32_t dim, int32_t axis,
                            double lo_cov, double hi_cov, int32_t k, const int64_t* idx_local, const double* dist,
                            const int64_t* local_to_global, const uint64_t* peer_result_blocks_h,
                            uint64_t* fails_per_asker /* device [n_ranks], may be NULL */, void* stream);
/* the deferred counts without a collective: every rank stores its counts (route overflow of its own queries; per
   asking rank, the queries it could not certify) into the same words of EVERY peer's flag block before the second
   barrier of the step; afterwards pomp_shard_worst_count reduces flags[rows][cols] to the largest column sum */
int pomp_shard_share_counts_p2p(const uint64_t* vals, int32_t ncols, const uint64_t* peer_flag_blocks_h,
                                int64_t dst_offset, int32_t n_ranks, void* stream);
int pomp_shard_worst_count(const int64_t* flags, int32_t rows, int32_t cols, int64_t* worst, void* stream);
int pomp_shard_unpack_p2p(const int64_t* rows, int32_t n_ranks, int64_t cap, int32_t k, const int64_t* perm,
                          const uint64_t* counters, int64_t* idx, double* dist, int64_t* defer_list, int64_t defer_cap,
                          void* stream);
int pomp_shard_scatter_deferred(const int64_t* merged_idx, const double* merged_dist, int32_t k,
                                const int64_t* defer_list, const uint64_t* defer_count, int64_t F, int64_t* idx,
                                double* dist, void* stream);

/* multi-GPU: merge G per-shard candidate lists [G][nq][k] (already carrying GLOBAL
   indices) into the k best per query by (dist, idx).  The exchange itself (NCCL all-to-all /
   all-gather over NVLink) is issued by the host binding through torch.distributed. */
int pomp_nn_merge_candidates(const int64_t* idx_in, const double* dist_in, int32_t n_shards, int64_t nq,
                             int32_t k, int64_t* idx_out, double* dist_out, int32_t* count_out,
                             void* stream);

/* ---------------------------------------------------------------- feasibility ---------- */
/* ConfigurationSpace.feasible for box-bounded product spaces with 2-D box / disc obstacles:
   BoxSet.contains (sets.py:177-182), Geometric2DCSpace.feasible (geometric.py:56-60),
   Circle.contains (:15-16), MultiConfigurationSpace.feasible (configurationspace.py:196-199). */
typedef struct {
  int32_t dim;
  int32_t obstacle_dim0;      /* the two coordinates the 2-D obstacles live in */
  int32_t obstacle_dim1;
  int32_t n_boxes;
  int32_t n_circles;
  int32_t _pad;
  double lo[POMP_MAX_DIM];    /* -inf / +inf for coordinates without a bound check */
  double hi[POMP_MAX_DIM];
  const double* boxes_h;      /* n_boxes x 4: xmin, ymin, xmax, ymax (closed)  */
  const double* circles_h;    /* n_circles x 3: cx, cy, r (closed disc)        */
} pomp_space_desc;

typedef struct pomp_space pomp_space;
int pomp_space_create(const pomp_space_desc* desc, int device, pomp_space** out);
int pomp_space_destroy(pomp_space* s);
/* CostControlSpace.setCostMax / in-place bound updates (costspace.py:26-30) */
int pomp_space_set_bounds(pomp_space* s, const double* lo_h, const double* hi_h);
/* "edge_overflow" (read-and-clear): 1 if an edge check since the last read asked for more than 1e7 samples
   (length / resolution; the reference, edgechecker.py:21-22, raises OverflowError or never returns there) --
   such an edge is reported infeasible and the *_h entry points return POMP_ERR_INVALID; device-pointer callers
   poll this after their stream has drained.  Also "n_boxes", "n_circles", "obstacle_grid_cells". */
int pomp_space_get_stat(pomp_space* s, const char* name, double* out);

int pomp_feasible(pomp_space* s, const double* x, int64_t n, uint8_t* ok, void* stream);
int pomp_feasible_h(pomp_space* s, const double* x_h, int64_t n, uint8_t* ok_h);

/* ---------------------------------------------------------------- edge checks ---------- */
/* EpsilonEdgeChecker.feasible (edgechecker.py:20-30) on LinearInterpolator edges
   (interpolators.py:28-35): a, b are n x dim. */
int pomp_edge_check(pomp_space* s, const double* a, const double* b, int64_t n, double resolution,
                    uint8_t* ok, void* stream);
int pomp_edge_check_h(pomp_space* s, const double* a_h, const double* b_h, int64_t n, double resolution,
                      uint8_t* ok_h);
/* same, edges given as index pairs ij[n_edges][2] into a roadmap node array nodes[*][dim] */
int pomp_edge_check_indexed(pomp_space* s, const double* nodes, const int32_t* ij, int64_t n_edges,
                            double resolution, uint8_t* ok, void* stream);
/* same with the node count known: ConfigurationSpace.feasible (geometric.py:56-60) is evaluated ONCE per node into
   node_ok[n_nodes] (device; returned to the caller) and the edges only gather the two flags and test their interior
   samples -- the result is the same conjunction (edgechecker.py:20-30); in a k-NN roadmap every node is an endpoint
   of ~2k edges */
int pomp_edge_check_indexed_nodes(pomp_space* s, const double* nodes, int64_t n_nodes, const int32_t* ij,
                                  int64_t n_edges, double resolution, uint8_t* node_ok, uint8_t* ok, void* stream);
/* PiecewiseLinearInterpolator / GeodesicInterpolator edges (interpolators.py:37-48,73-102):
   path p has waypoints [path_offsets[p], path_offsets[p+1]) of `waypoints` (row-major x dim);
   length / interpolation follow `geodesic` (EUCLIDEAN or GEODESIC_PRODUCT). */
int pomp_edge_check_path(pomp_space* s, const pomp_metric_desc* geodesic, const double* waypoints,
                         const int32_t* path_offsets, int64_t n_paths, double resolution, uint8_t* ok,
                         void* stream);
int pomp_edge_check_path_h(pomp_space* s, const pomp_metric_desc* geodesic, const double* waypoints_h,
                           const int32_t* path_offsets_h, int64_t n_paths, double resolution,
                           uint8_t* ok_h);

/* ---------------------------------------------------------------- propagation ---------- */
typedef enum {
  POMP_DYN_ADAPTOR = 0,    /* ControlSpaceAdaptor.nextState: x' = u           (controlspace.py:78-79) */
  POMP_DYN_CV = 1,         /* CVControlSpace.nextState closed form            (statespace.py:62-70)   */
  POMP_DYN_CV_EULER = 2,   /* CVControlSpace.trajectory()[-1] stepwise        (statespace.py:47-61)   */
  POMP_DYN_DUBINS = 3,     /* DubinsCarSpace.nextState                        (dubins.py:107-125)     */
  POMP_DYN_PENDULUM = 4,   /* LambdaKinodynamicSpace Euler + Pendulum.derivative (controlspace.py:249-267, pendulum.py:152-154) */
  POMP_DYN_LTI = 5,        /* LTIControlSpace: A x + B u                      (controlspace.py:101-102) */
  POMP_DYN_FLAPPY = 6      /* FlappyControlSpace.eval(x,u,1)                  (flappy.py:20-31)        */
} pomp_dyn_kind;

typedef enum {
  POMP_COST_NONE = 0,       /* plain ControlSpace                                                     */
  POMP_COST_TIME = 1,       /* CostControlSpace + TimeObjectiveFunction: c += u[0]       (objectives.py) */
  POMP_COST_ABS_U0 = 2,     /* DubinsCarDistanceObjectiveFunction(1): c += |u[0]|       (dubins.py:131-135) */
  POMP_COST_PATH_LENGTH = 3 /* PathLengthObjectiveFunction on an adaptor: c += euclid(x,u)             */
} pomp_cost_kind;

typedef struct {
  int32_t kind;      /* pomp_dyn_kind */
  int32_t x_dim;     /* base state dimension (without cost coordinate) */
  int32_t u_dim;     /* control dimension (u[0] = duration for the integrated kinds) */
  int32_t cost_kind; /* pomp_cost_kind: when != NONE states carry a trailing cost coordinate (costspace.py:35-38) */
  double dt;         /* integration step (KinodynamicSpace.dt) */
  double p[8];       /* PENDULUM: g, m, L ; FLAPPY: v_x, gravity, thrust */
  double A[POMP_MAX_DIM * POMP_MAX_DIM]; /* LTI, row-major x_dim x x_dim */
  double B[POMP_MAX_DIM * POMP_MAX_DIM]; /* LTI, row-major x_dim x u_dim */
} pomp_dyn_desc;

/* ControlSpace.nextState for n (x,u) pairs; x: n x X, u: n x u_dim, xnext: n x X with
   X = x_dim (+1 with a cost objective). */
int pomp_next_state(const pomp_dyn_desc* dyn, const double* x, const double* u, int64_t n, double* xnext,
                    void* stream);
int pomp_next_state_h(const pomp_dyn_desc* dyn, const double* x_h, const double* u_h, int64_t n,
                      double* xnext_h);
/* RandomControlSelector.select (kinodynamicplanner.py:66-84), fused nextState + metric + strict-<
   arg-min (first minimum wins): for each of B parents, S candidate controls u[B][S][u_dim];
   valid[B][S] = U.contains(u).  best[b] = -1 when no valid candidate has finite distance. */
int pomp_select_control(const pomp_dyn_desc* dyn, const pomp_metric_desc* metric, const double* x,
                        const double* xdes, const double* u, const uint8_t* valid, int64_t B, int32_t S,
                        int32_t* best, double* xnext_best, double* dist_best, void* stream);
int pomp_select_control_h(const pomp_dyn_desc* dyn, const pomp_metric_desc* metric, const double* x_h,
                          const double* xdes_h, const double* u_h, const uint8_t* valid_h, int64_t B,
                          int32_t S, int32_t* best_h, double* xnext_best_h, double* dist_best_h);

/* EpsilonEdgeChecker.feasible on controlSpace.interpolator(x,u) without materialising the curve on
   the host: ADAPTOR -> LinearInterpolator(x,u); DUBINS -> DubinsCarInterpolator (dubins.py:62-73:
   length |u0|, eval(s) = nextState(x,[u0*s,u1])); PENDULUM / CV_EULER ->
   PiecewiseLinearInterpolator(trajectory(x,u)) under `geodesic` (controlspace.py:149-172,249-271).
   x: n x x_dim BASE states (CostEdgeChecker checks components[0], kinodynamicplanner.py:895-896). */
int pomp_edge_check_control(pomp_space* s, const pomp_dyn_desc* dyn, const pomp_metric_desc* geodesic,
                            const double* x, const double* u, int64_t n, double resolution, uint8_t* ok,
                            void* stream);
int pomp_edge_check_control_h(pomp_space* s, const pomp_dyn_desc* dyn, const pomp_metric_desc* geodesic,
                              const double* x_h, const double* u_h, int64_t n, double resolution,
                              uint8_t* ok_h);

/* ---------------------------------------------------------------- fused RRT extend ------ */
/* One kinematic RRT.expand step (kinodynamicplanner.py:354-411 with KinematicControlSelector
   :95-100, LinearInterpolator, EpsilonEdgeChecker) in ONE launch: [feasible(xrand)] ->
   nearest(xrand) -> steer to at most max_step -> sampled edge check -> append to the index on success.
   The host supplies the already drawn xrand (Python's RNG stays on the host); check_sample != 0
   also runs RRT.expand's cspace.feasible(xrand) test (:362) on the device.
   Outputs (host): parent index (-1: no admissible node, -2: xrand infeasible), the new point
   (dim doubles), the edge length |parent - xnew| (PathLengthObjectiveFunction.incremental) and
   whether the node was appended.  The metric of `h` must be EUCLIDEAN. */
int pomp_rrt_extend_h(pomp_nn* h, pomp_space* s, const double* xrand_h, int check_sample, double max_step,
                      double resolution, int64_t* parent_h, double* xnew_h, double* edge_len_h,
                      int32_t* added_h);

/* ---------------------------------------------------------------- lock-step trials ------ */
/* Gaussian kernel density of EST: density[i] = sum over the nodes p with metric(p, q_i) <= radius
 * (inclusive != 0) or < radius of exp(-(metric(p, q_i)/sigma)^2); count[i] (nullable) = number of such nodes.  Replaces
 * ESTKinodynamicPlanner.density (pomp/planners/kinodynamicplanner.py:688-693: neighbors(x, 4*radius), then the Python
 * sum) and the neighbour loop of onAddNode (:676-686).  The hits are folded on the fly, nothing is materialised.  The
 * summation order is not the reference's (which itself depends on the NN method): agreement to ~1e-13 relative. */
int pomp_nn_density(pomp_nn* h, const double* q, int64_t nq, double radius, double sigma, int inclusive,
                    double* density, int32_t* count, void* stream);
int pomp_nn_density_h(pomp_nn* h, const double* q_h, int64_t nq, double radius, double sigma, int inclusive,
                      double* density_h, int32_t* count_h);

/* RRT* rewire batch.  Replaces, for one new node x, the knearest call of RRTStar.rewire
 * (pomp/planners/rrtstarplanner.py:135-139) and the up-to-2k EpsilonEdgeChecker.feasible calls on straight edges that
 * follow it (:150-152 neighbour -> x, :177-179 x -> neighbour) by one call: idx/dist/count as pomp_nn_knearest_h for
 * the single query x; ok_in[j] = edge (node idx[j] -> x) feasible, ok_out[j] = edge (x -> node idx[j]) feasible, both
 * with the reference's sampling schedule (edgechecker.py:20-30).  Entries j >= count are unspecified.  Host pointers. */
int pomp_rewire_candidates_h(pomp_nn* h, pomp_space* s, const double* x_h, int32_t k, double resolution,
                             int64_t* idx_h, double* dist_h, int32_t* count_h, uint8_t* ok_in_h, uint8_t* ok_out_h);

/* T independent kinematic RRTs (main.py runs 10 trials per configuration, pomp/planners/test.py:8-52;
   BASELINE config 5 asks for thousands) advanced by one RRT.expand EACH in one launch: one block per
   trial over a device arena [T][capacity][dim].  Per trial the semantics are exactly those of
   pomp_rrt_extend_h; reset_h[t] != 0 restarts trial t from root_h first (RRT.reset,
   kinodynamicplanner.py:312-320); active_h[t] == 0 skips it (parent -3).  The host keeps one RNG per
   trial, so every tree equals the single-trial one.  POMP_ERR_CAPACITY if a tree is full. */
typedef struct pomp_forest pomp_forest;
int pomp_forest_create(int dim, int n_trials, int64_t capacity_per_trial, int device, pomp_forest** out);
int pomp_forest_destroy(pomp_forest* f);
int64_t pomp_forest_size(const pomp_forest* f, int trial);
int pomp_forest_extend_h(pomp_forest* f, pomp_space* s, const double* xrand_h, const double* root_h,
                         const uint8_t* active_h, const uint8_t* reset_h, int check_sample, double max_step,
                         double resolution, int64_t* parent_h, double* xnew_h, double* edge_len_h,
                         int32_t* added_h);

/* n_iters (1..32) expansions per trial in ONE launch: xrand_h is [T][n_iters][dim], outputs are [T][n_iters](...).
   Trial t executes its samples in order exactly as n_iters successive pomp_forest_extend_h calls would, except that
   it stops after the expansion whose new node lies within goal_radius of goal_h (NULL: never) -- where
   RepeatedRRT.planMore returns (kinodynamicplanner.py:1291-1312).  n_done_h[t] = expansions executed (0 for an
   inactive trial); only the first n_done_h[t] output records of a trial are written.  The planner loop is bound by
   launch + synchronisation latency; RRT's samples do not depend on the tree, so they can be drawn ahead. */
int pomp_forest_extend_multi_h(pomp_forest* f, pomp_space* s, int32_t n_iters, const double* xrand_h,
                               const double* root_h, const uint8_t* active_h, const uint8_t* reset_h, int check_sample,
                               double max_step, double resolution, const double* goal_h, double goal_radius,
                               int32_t* n_done_h, int64_t* parent_h, double* xnew_h, double* edge_len_h,
                               int32_t* added_h);

/* ---------------------------------------------------------------- device-resident roadmap ---------------- */
/* k-nearest-neighbour roadmap over the node set of `nn` in space `s` (BASELINE config 3), every step a call of the
   path: ConfigurationSpace.feasible per node (geometric.py:56-60) -> node_ok; NearestNeighbors.knearest(x, k+1) per
   node (nearestneighbors.py:98-116; the node itself dropped); EpsilonEdgeChecker.feasible per candidate edge
   (edgechecker.py:20-30); stable compaction of the feasible edges (node order, then neighbour rank) as index pairs
   -- the (V, E) of TreePlanner.getRoadmap (kinodynamicplanner.py:211-238) without leaving the device.
   undirected != 0: of a mutual pair only (min, max) is kept.  *n_edges_h may exceed cap (nothing is written beyond
   cap; call again with more room).  Device pointers + stream in the first form, host buffers in the _h form. */
int pomp_roadmap_build(pomp_nn* nn, pomp_space* s, int32_t k, double resolution, int32_t undirected,
                       uint8_t* node_ok, int32_t* edges_out, int64_t cap, int64_t* n_edges_h,
                       int64_t* n_candidates_h, void* stream);
int pomp_roadmap_build_h(pomp_nn* nn, pomp_space* s, int32_t k, double resolution, int32_t undirected,
                         uint8_t* node_ok_h, int32_t* edges_h, int64_t cap, int64_t* n_edges_h,
                         int64_t* n_candidates_h);

/* ---------------------------------------------------------------- EST projection-hash density ------------ */
/* ESTWithProjections (kinodynamicplanner.py:746-868): nodes are hashed per projection basis b (<= 4 coordinates
   basis_idx[b][i] with scales basis_scale[b][i], rows padded to 4) to the cell key (int(x[k]*v))_k; density(x) =
   sum_b (#nodes with the key of x in basis b) * basis_weight[b] / n_bases (:834-856; weight = pow(level, 1)).
   add = onAddNode (:817-832); the node coordinates live on the device, a warp per query counts the matches. */
typedef struct pomp_est pomp_est;
int pomp_est_create(int32_t dim, int32_t n_bases, const int32_t* basis_dim_h, const int32_t* basis_idx_h,
                    const double* basis_scale_h, const double* basis_weight_h, int device, pomp_est** out);
int pomp_est_destroy(pomp_est* e);
int pomp_est_reset(pomp_est* e);
int64_t pomp_est_size(const pomp_est* e);
int pomp_est_add_h(pomp_est* e, const double* pts_h, int64_t n);
int pomp_est_density(pomp_est* e, const double* q, int64_t nq, double* out, void* stream);
int pomp_est_density_h(pomp_est* e, const double* q_h, int64_t nq, double* out_h);

/* ---------------------------------------------------------------- lock-step sst* forest ------------------ */
/* n_trials independent StableSparseRRT trials (rrtstarplanner.py:203-410 under the restart schedule of
   StableSparseRRTStar :454-471), one block per trial, n_iters planner iterations per call.  The caller supplies, per
   trial and iteration, what the reference draws from Python's `random` (sample xrand[x_dim], control u[u_dim] =
   (duration, base control), valid = controlSet.contains(u)) and the per-iteration selection / witness radii and
   restart flags (restart[m] = 1: the tree is reset to the root BEFORE iteration m; restart_now: before the first).
   On the device: pickNode, the stepwise trajectory of `dyn` as a piecewise-linear edge under `geodesic`, the sampled
   edge check against `space`, costs (+ u[0]), witness bookkeeping, pruning, goal test (geodesic distance to
   goal_center <= goal_radius; goal_center NULL = none).  Outputs per trial: best goal cost so far (+inf = none),
   nodes created since the last restart, iterations in which no active node existed (must be 0 for the pre-drawn
   random stream to stay in step).  POMP_ERR_CAPACITY when an arena overflows.
   half_sq_last_h[trial][iter] / half_sq_full (constant-acceleration steps, statespace.py:55-59): the caller's
   0.5*dt**2 for the last (partial) and for a full integration step -- Python's ** is libm pow, which differs from
   dt*dt in ~0.1 % of inputs, and one ulp in a node coordinate is a different tree; NULL: computed on the device. */
typedef struct pomp_sst pomp_sst;
int pomp_sst_create(int32_t x_dim, int32_t u_dim, int32_t n_trials, int32_t node_capacity, int32_t witness_capacity,
                    const pomp_dyn_desc* dyn, const pomp_metric_desc* geodesic, int device, pomp_sst** out);
int pomp_sst_destroy(pomp_sst* f);
int pomp_sst_reset_h(pomp_sst* f, const double* root_h);
int pomp_sst_step_h(pomp_sst* f, pomp_space* space, int32_t n_iters, const double* xrand_h, const double* u_h,
                    const uint8_t* valid_h, const double* sel_radius_h, const double* wit_radius_h,
                    const uint8_t* restart_h, int32_t restart_now, const double* goal_center_h, double goal_radius,
                    double resolution, const double* half_sq_last_h, double half_sq_full, double* best_cost_h,
                    int32_t* n_nodes_h, int32_t* n_none_h);
/* node arrays of one trial in creation order (any output may be NULL): state, parent (-1: root / removed), alive
   (still in the tree), cost, active; *count_h = nodes created since the last restart */
int pomp_sst_get_trial_h(pomp_sst* f, int32_t trial, double* x_h, int32_t* parent_h, uint8_t* alive_h, double* cost_h,
                         uint8_t* active_h, int32_t* count_h);

#ifdef __cplusplus
}
#endif
#endif /* POMP_B200_H */
