/*
 * pomp_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement, in plain C, of the reference
 * algorithms on the hot path of krishauser/pyOptimalMotionPlanning (pure Python, /root/reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (pyoptimalmotionplanning_b200) never does.
 *
 * Parity pin: the reference ships no tests or golden vectors for this path.  This restatement
 * is pinned instead against outputs of the unmodified reference itself, generated in the build
 * container by oracle/make_golden.py (fixtures in tests/golden/ JSON files) and checked by
 * tests/test_oracle_golden.py.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math -fno-builtin-pow (see oracle/Makefile).
 * -ffp-contract=off matters: the reference rounds after every operation (CPython floats).
 *
 * Every function cites the reference file:line it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/pomp_b200.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ scalar helpers ------ */

/* CPython float.__mod__ (Objects/floatobject.c float_rem): fmod, then shift into the divisor's
   sign; used by so2.normalize (pomp/klampt/so2.py:23-28). */
static double py_fmod(double a, double m) {
  double r = fmod(a, m);
  if (r != 0.0) {
    if ((m < 0.0) != (r < 0.0)) r += m;
  } else {
    r = copysign(0.0, m);
  }
  return r;
}

static const double ORC_PI = 3.141592653589793;     /* math.pi */
static const double ORC_TWOPI = 6.283185307179586;  /* math.pi*2 */

/* pomp/klampt/so2.py:23-28 */
static double so2_normalize(double a) {
  double res = py_fmod(a, ORC_TWOPI);
  if (res < 0) return res + ORC_TWOPI;
  return res;
}

/* pomp/klampt/so2.py:30-37 */
static double so2_diff(double a, double b) {
  double d = so2_normalize(a) - so2_normalize(b);
  if (d < -ORC_PI) return d + ORC_TWOPI;
  if (d > ORC_PI) return d - ORC_TWOPI;
  return d;
}

/* pomp/klampt/vectorops.py:94-102 (distanceSquared + sqrt), sum starts at int 0 */
static double euclid(const double* a, const double* b, int d) {
  double s = 0.0;
  for (int i = 0; i < d; i++) s = s + (a[i] - b[i]) * (a[i] - b[i]);
  return sqrt(s);
}


/* CPython >= 3.12 builtin sum() over floats (Python/bltinmodule.c, builtin_sum_impl): Neumaier
   compensated summation, compensation added once at the end.  The reference's L1Metric
   (metric.py:10-11) and WeightedEuclideanMetric (:19-21) go through it; the golden vectors were
   produced under CPython 3.12.3, so the oracle follows that interpreter. */
typedef struct { double s, c; } py_sum_t;
static void py_sum_add(py_sum_t* a, double x) {
  double t = a->s + x;
  if (fabs(a->s) >= fabs(x)) a->c += (a->s - t) + x;
  else a->c += (x - t) + a->s;
  a->s = t;
}
static double py_sum_result(const py_sum_t* a) {
  if (a->c != 0.0 && isfinite(a->c)) return a->s + a->c;
  return a->s;
}

/* ------------------------------------------------------------------ metrics ------------- */

/* base metric without the cost wrapper, on the first `d` coordinates */
static double base_metric(const pomp_metric_desc* m, const double* a, const double* b, int d) {
  switch (m->kind) {
    case POMP_METRIC_EUCLIDEAN: /* pomp/spaces/metric.py:4-5 */
      return euclid(a, b, d);
    case POMP_METRIC_WEIGHTED_EUCLIDEAN: { /* metric.py:19-21: sqrt(sum(w*(ai-bi)**2)); ** is libm pow */
      py_sum_t s = {0.0, 0.0};
      for (int i = 0; i < d; i++) py_sum_add(&s, m->dim_weight[i] * pow(a[i] - b[i], 2.0));
      return sqrt(py_sum_result(&s));
    }
    case POMP_METRIC_SCALED_EUCLIDEAN: /* metric.py:22-23 */
      return m->scale * euclid(a, b, d);
    case POMP_METRIC_L1: { /* metric.py:10-11 */
      py_sum_t s = {0.0, 0.0};
      for (int i = 0; i < d; i++) py_sum_add(&s, fabs(a[i] - b[i]));
      return py_sum_result(&s);
    }
    case POMP_METRIC_LINF: { /* metric.py:13-14 */
      double s = fabs(a[0] - b[0]);
      for (int i = 1; i < d; i++) {
        double v = fabs(a[i] - b[i]);
        if (v > s) s = v;
      }
      return s;
    }
    case POMP_METRIC_GEODESIC_PRODUCT: { /* pomp/spaces/geodesicspace.py:45-52 */
      double res = 0.0;
      int i = 0;
      for (int c = 0; c < m->n_comp; c++) {
        int cd = m->comp_dim[c];
        double dc;
        if (m->comp_kind[c] == POMP_COMP_SO2)
          dc = fabs(so2_diff(a[i], b[i])); /* pomp/spaces/so2space.py:9-10 */
        else
          dc = euclid(a + i, b + i, cd); /* geodesicspace.py:7-8 */
        res += pow(dc, 2.0) * m->comp_weight[c];
        i += cd;
      }
      return sqrt(res);
    }
  }
  return NAN;
}

/* metric(a,b): a = tree node, b = query.  CostMetric.__call__: kinodynamicplanner.py:905-913 */
ORC_API double orc_metric(const pomp_metric_desc* m, const double* a, const double* b) {
  if (!m->has_cost) return base_metric(m, a, b, m->dim);
  int d = m->dim - 1;
  double dist = base_metric(m, a, b, d);
  if (m->cost_weight_set) {
    if (m->cost_max_set && (a[d] > m->cost_max || b[d] > m->cost_max)) return INFINITY;
    double dc = a[d] - b[d];
    dist += (dc > 0.0 ? dc : 0.0) * m->cost_weight; /* max(a[-1]-b[-1],0.0)*costWeight */
  }
  return dist;
}

/* ------------------------------------------------------------------ brute-force NN ------ */
/* mask[i] != 0  <=>  the reference's filter(point,data) is True (reject) or the point was removed. */

/* NearestNeighbors.nearest, brute-force branch: nearestneighbors.py:88-96 (strict <, first wins) */
ORC_API void orc_nn_nearest(const pomp_metric_desc* m, const double* pts, int64_t n, const uint8_t* mask,
                            const double* q, int64_t nq, int64_t* idx, double* dist) {
  int D = m->dim;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t j = 0; j < nq; j++) {
    int64_t res = -1;
    double dbest = INFINITY;
    for (int64_t i = 0; i < n; i++) {
      double d = orc_metric(m, pts + i * D, q + j * D);
      if (d < dbest && !(mask && mask[i])) {
        res = i;
        dbest = d;
      }
    }
    idx[j] = res;
    dist[j] = dbest;
  }
}

/* KNearestResult (pomp/structures/knn.py:7-30) */
typedef struct {
  int k, imin, imax;
  int64_t* items;
  double* distances;
} knn_result;

static void knn_init(knn_result* r, int k, int64_t* items, double* distances) {
  r->k = k;
  r->imin = r->imax = 0;
  r->items = items;
  r->distances = distances;
  for (int i = 0; i < k; i++) {
    items[i] = -1;
    distances[i] = INFINITY;
  }
}

/* knn.py:14-23 */
static void knn_tryadd(knn_result* r, int64_t item, double distance) {
  if (distance < r->distances[r->imax]) {
    r->distances[r->imax] = distance;
    r->items[r->imax] = item;
    if (distance < r->distances[r->imin]) r->imin = r->imax;
    for (int i = 0; i < r->k; i++)
      if (r->distances[i] > r->distances[r->imax]) r->imax = i;
  }
}

/* knn.py:28-30: sorted by (distance, item); with unique insertion indices standing in for the
   (point,data) item this is the (distance, index) order.  Returns the count. */
static int knn_sorted(knn_result* r) {
  int cnt = 0;
  for (int i = 0; i < r->k; i++)
    if (r->distances[i] != INFINITY) {
      r->distances[cnt] = r->distances[i];
      r->items[cnt] = r->items[i];
      cnt++;
    }
  for (int i = 1; i < cnt; i++) { /* insertion sort, k is small */
    double d = r->distances[i];
    int64_t it = r->items[i];
    int j = i - 1;
    while (j >= 0 && (r->distances[j] > d || (r->distances[j] == d && r->items[j] > it))) {
      r->distances[j + 1] = r->distances[j];
      r->items[j + 1] = r->items[j];
      j--;
    }
    r->distances[j + 1] = d;
    r->items[j + 1] = it;
  }
  for (int i = cnt; i < r->k; i++) {
    r->distances[i] = INFINITY;
    r->items[i] = -1;
  }
  return cnt;
}

/* NearestNeighbors.knearest, brute-force branch as intended: nearestneighbors.py:110-116
   (the shipped call has its tryadd arguments swapped -- SURVEY appendix B#1 -- the kd-tree branch
   kdtree.py:299-303 shows the intended (item, distance) order). */
ORC_API void orc_nn_knearest(const pomp_metric_desc* m, const double* pts, int64_t n, const uint8_t* mask,
                             const double* q, int64_t nq, int k, int64_t* idx, double* dist,
                             int32_t* count) {
  int D = m->dim;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t j = 0; j < nq; j++) {
    knn_result r;
    knn_init(&r, k, idx + j * k, dist + j * k);
    for (int64_t i = 0; i < n; i++) {
      if (mask && mask[i]) continue;
      knn_tryadd(&r, i, orc_metric(m, pts + i * D, q + j * D));
    }
    count[j] = knn_sorted(&r);
  }
}

/* NearestNeighbors.neighbors: nearestneighbors.py:134-141 (d < radius) / kdtree.py:347 (d <= rad).
   Results in insertion order.  Two-pass CSR; returns total hits (fills idx only if it fits). */
ORC_API int64_t orc_nn_neighbors(const pomp_metric_desc* m, const double* pts, int64_t n,
                                 const uint8_t* mask, const double* q, int64_t nq, double radius,
                                 int inclusive, int64_t* offsets, int64_t* idx, int64_t cap) {
  int D = m->dim;
  offsets[0] = 0;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t j = 0; j < nq; j++) {
    int64_t c = 0;
    for (int64_t i = 0; i < n; i++) {
      if (mask && mask[i]) continue;
      double d = orc_metric(m, pts + i * D, q + j * D);
      if (inclusive ? (d <= radius) : (d < radius)) c++;
    }
    offsets[j + 1] = c;
  }
  for (int64_t j = 0; j < nq; j++) offsets[j + 1] += offsets[j];
  if (offsets[nq] > cap) return offsets[nq];
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t j = 0; j < nq; j++) {
    int64_t o = offsets[j];
    for (int64_t i = 0; i < n; i++) {
      if (mask && mask[i]) continue;
      double d = orc_metric(m, pts + i * D, q + j * D);
      if (inclusive ? (d <= radius) : (d < radius)) idx[o++] = i;
    }
  }
  return offsets[nq];
}

/* ------------------------------------------------------------------ kd-tree ------------- */
/* Restatement of pomp/structures/kdtree.py (the reference's DEFAULT nearestNeighborMethod):
   midpoint-of-range splits, leaf cap 10, metric-generic plane distances.  Used as the CPU
   baseline and to cross-check brute force on Euclidean metrics. */

typedef struct kd_node {
  int32_t* pts; /* leaf: indices of points (kdtree.py:23 Node.points) */
  int32_t npts, cap;
  int32_t splitdim, depth;
  int has_split;
  double splitvalue;
  struct kd_node *left, *right;
} kd_node;

typedef struct {
  pomp_metric_desc metric;
  const double* pts; /* borrowed, n x D */
  int D;
  int max_pts_per_node; /* kdtree.py:78 */
  int max_depth;
  int64_t num_nodes;
  kd_node* root;
} orc_kdtree;

static kd_node* kd_new_node(int splitdim) {
  kd_node* nd = (kd_node*)calloc(1, sizeof(kd_node));
  nd->splitdim = splitdim;
  return nd;
}

static void kd_push(kd_node* nd, int32_t i) {
  if (nd->npts == nd->cap) {
    nd->cap = nd->cap ? nd->cap * 2 : 12;
    nd->pts = (int32_t*)realloc(nd->pts, sizeof(int32_t) * (size_t)nd->cap);
  }
  nd->pts[nd->npts++] = i;
}

static void kd_free_node(kd_node* nd) {
  if (!nd) return;
  kd_free_node(nd->left);
  kd_free_node(nd->right);
  free(nd->pts);
  free(nd);
}

/* kdtree.py:99-160 recursive_split */
static int kd_recursive_split(orc_kdtree* t, kd_node* nd, int force, int optimize) {
  if (!force && nd->npts <= t->max_pts_per_node) return 0;
  if (nd->npts == 0) return 0;
  int d = t->D;
  const double* P = t->pts;
  double vmin = 0, vmax = 0;
  if (!optimize) {
    for (int i = 0; i < d; i++) {
      vmin = INFINITY;
      vmax = -INFINITY;
      for (int j = 0; j < nd->npts; j++) {
        double v = P[(int64_t)nd->pts[j] * d + nd->splitdim];
        if (v < vmin) vmin = v;
        if (v > vmax) vmax = v;
      }
      if (vmin != vmax) break;
      nd->splitdim = (nd->splitdim + 1) % d;
    }
  } else {
    double rmin = 0, rmax = 0;
    int dimmax = 0;
    for (int i = 0; i < d; i++) {
      vmin = INFINITY;
      vmax = -INFINITY;
      for (int j = 0; j < nd->npts; j++) {
        double v = P[(int64_t)nd->pts[j] * d + i];
        if (v < vmin) vmin = v;
        if (v > vmax) vmax = v;
      }
      if (vmax - vmin > rmax - rmin) {
        rmin = vmin;
        rmax = vmax;
        dimmax = i;
      }
    }
    nd->splitdim = dimmax;
    vmin = rmin;
    vmax = rmax;
  }
  if (vmin == vmax) return 0;
  double sv = (vmin + vmax) * 0.5;
  kd_node* L = kd_new_node((nd->splitdim + 1) % d);
  kd_node* R = kd_new_node((nd->splitdim + 1) % d);
  for (int j = 0; j < nd->npts; j++) {
    /* partitionFn: copysign(1, x[d]-value) < 0 -> left (kdtree.py:49,145-147) */
    double diff = P[(int64_t)nd->pts[j] * d + nd->splitdim] - sv;
    if (signbit(diff)) kd_push(L, nd->pts[j]);
    else kd_push(R, nd->pts[j]);
  }
  if (L->npts == 0 || R->npts == 0) {
    kd_free_node(L);
    kd_free_node(R);
    return 0;
  }
  nd->has_split = 1;
  nd->splitvalue = sv;
  nd->left = L;
  nd->right = R;
  L->depth = R->depth = nd->depth + 1;
  t->num_nodes += 2;
  if (nd->depth + 1 > t->max_depth) t->max_depth = nd->depth + 1;
  int d1 = kd_recursive_split(t, L, 0, optimize);
  int d2 = kd_recursive_split(t, R, 0, optimize);
  free(nd->pts);
  nd->pts = NULL;
  nd->npts = nd->cap = 0;
  return 1 + (d1 > d2 ? d1 : d2);
}

ORC_API orc_kdtree* orc_kd_create(const pomp_metric_desc* m, const double* pts) {
  orc_kdtree* t = (orc_kdtree*)calloc(1, sizeof(orc_kdtree));
  t->metric = *m;
  t->pts = pts;
  t->D = m->dim;
  t->max_pts_per_node = 10;
  return t;
}

ORC_API void orc_kd_destroy(orc_kdtree* t) {
  if (!t) return;
  kd_free_node(t->root);
  free(t);
}

/* KDTree.set (kdtree.py:91-97): bulk build over points 0..n-1 */
ORC_API void orc_kd_set(orc_kdtree* t, const double* pts, int64_t n) {
  kd_free_node(t->root);
  t->pts = pts;
  t->max_depth = 0;
  t->num_nodes = 1;
  t->root = kd_new_node(0);
  for (int64_t i = 0; i < n; i++) kd_push(t->root, (int32_t)i);
  kd_recursive_split(t, t->root, 0, 1);
}

/* KDTree.add (kdtree.py:162-174): insert point index i (its coordinates already live in pts) */
ORC_API void orc_kd_add(orc_kdtree* t, const double* pts, int64_t i) {
  t->pts = pts;
  if (!t->root) {
    t->root = kd_new_node(0);
    kd_push(t->root, (int32_t)i);
    t->max_depth = 0;
    t->num_nodes = 1;
    return;
  }
  kd_node* nd = t->root;
  const double* x = pts + i * t->D;
  while (nd->has_split) nd = signbit(x[nd->splitdim] - nd->splitvalue) ? nd->left : nd->right;
  kd_push(nd, (int32_t)i);
  kd_recursive_split(t, nd, 0, 0);
}

ORC_API int64_t orc_kd_stats(const orc_kdtree* t, int which) {
  return which == 0 ? t->num_nodes : t->max_depth;
}

/* minDistanceFn (kdtree.py:51-72): |x[d]-value| for euclideanMetric, else metric(x with x[d]:=value, x) */
static double kd_plane_dist(const orc_kdtree* t, int dim, double value, const double* x) {
  if (t->metric.kind == POMP_METRIC_EUCLIDEAN && !t->metric.has_cost) return fabs(x[dim] - value);
  double tmp[POMP_MAX_DIM];
  memcpy(tmp, x, sizeof(double) * (size_t)t->D);
  tmp[dim] = value;
  return orc_metric(&t->metric, tmp, x);
}

/* kdtree.py:237-284 */
static void kd_nearest_rec(const orc_kdtree* t, const kd_node* nd, const double* x, const uint8_t* mask,
                           double dmin, int32_t* closest, double* dout) {
  if (!nd->has_split) {
    int32_t c = -1;
    for (int j = 0; j < nd->npts; j++) {
      int32_t p = nd->pts[j];
      if (mask && mask[p]) continue;
      double d = orc_metric(&t->metric, t->pts + (int64_t)p * t->D, x);
      if (d < dmin) {
        c = p;
        dmin = d;
      }
    }
    *closest = c;
    *dout = dmin;
    return;
  }
  double dhi = 0, dlo = 0;
  if (signbit(x[nd->splitdim] - nd->splitvalue)) dhi = kd_plane_dist(t, nd->splitdim, nd->splitvalue, x);
  else dlo = kd_plane_dist(t, nd->splitdim, nd->splitvalue, x);
  if (dhi > dmin) {
    kd_nearest_rec(t, nd->left, x, mask, dmin, closest, dout);
    return;
  } else if (dlo > dmin) {
    kd_nearest_rec(t, nd->right, x, mask, dmin, closest, dout);
    return;
  }
  const kd_node *first = nd->left, *second = nd->right;
  if (dlo > dhi) {
    first = nd->right;
    second = nd->left;
    double tmp = dlo;
    dlo = dhi;
    dhi = tmp;
  }
  int32_t best = -1, c;
  double d;
  kd_nearest_rec(t, first, x, mask, dmin, &c, &d);
  if (d < dmin) {
    best = c;
    dmin = d;
  }
  if (dhi < dmin) {
    kd_nearest_rec(t, second, x, mask, dmin, &c, &d);
    if (d < dmin) {
      best = c;
      dmin = d;
    }
  }
  *closest = best;
  *dout = dmin;
}

/* KDTree.nearest (kdtree.py:286-293) */
ORC_API void orc_kd_nearest(const orc_kdtree* t, const uint8_t* mask, const double* q, int64_t nq,
                            int64_t* idx, double* dist) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t j = 0; j < nq; j++) {
    int32_t c = -1;
    double d = INFINITY;
    if (t->root) kd_nearest_rec(t, t->root, q + j * t->D, mask, INFINITY, &c, &d);
    idx[j] = c;
    dist[j] = d;
  }
}

/* kdtree.py:295-331 */
static void kd_knearest_rec(const orc_kdtree* t, const kd_node* nd, const double* x, const uint8_t* mask,
                            knn_result* res) {
  if (!nd->has_split) {
    for (int j = 0; j < nd->npts; j++) {
      int32_t p = nd->pts[j];
      if (mask && mask[p]) continue;
      knn_tryadd(res, p, orc_metric(&t->metric, t->pts + (int64_t)p * t->D, x));
    }
    return;
  }
  double dhi = 0, dlo = 0;
  if (signbit(x[nd->splitdim] - nd->splitvalue)) dhi = kd_plane_dist(t, nd->splitdim, nd->splitvalue, x);
  else dlo = kd_plane_dist(t, nd->splitdim, nd->splitvalue, x);
  if (dhi > res->distances[res->imax]) {
    kd_knearest_rec(t, nd->left, x, mask, res);
  } else if (dlo > res->distances[res->imax]) {
    kd_knearest_rec(t, nd->right, x, mask, res);
  } else {
    const kd_node *first = nd->left, *second = nd->right;
    if (dlo > dhi) {
      first = nd->right;
      second = nd->left;
      double tmp = dlo;
      dlo = dhi;
      dhi = tmp;
    }
    kd_knearest_rec(t, first, x, mask, res);
    if (dhi < res->distances[res->imax]) kd_knearest_rec(t, second, x, mask, res);
  }
}

/* KDTree.knearest (kdtree.py:334-340) */
ORC_API void orc_kd_knearest(const orc_kdtree* t, const uint8_t* mask, const double* q, int64_t nq, int k,
                             int64_t* idx, double* dist, int32_t* count) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t j = 0; j < nq; j++) {
    knn_result r;
    knn_init(&r, k, idx + j * k, dist + j * k);
    if (t->root) kd_knearest_rec(t, t->root, q + j * t->D, mask, &r);
    count[j] = knn_sorted(&r);
  }
}

/* kdtree.py:342-363; hits are appended in traversal order (right subtree first) */
static void kd_neighbors_rec(const orc_kdtree* t, const kd_node* nd, const double* x, double rad,
                             int64_t* out, int64_t* cnt, int64_t cap) {
  if (!nd->has_split) {
    for (int j = 0; j < nd->npts; j++) {
      int32_t p = nd->pts[j];
      double d = orc_metric(&t->metric, t->pts + (int64_t)p * t->D, x);
      if (d <= rad) {
        if (*cnt < cap) out[*cnt] = p;
        (*cnt)++;
      }
    }
    return;
  }
  double dhi = 0, dlo = 0;
  if (signbit(x[nd->splitdim] - nd->splitvalue)) dhi = kd_plane_dist(t, nd->splitdim, nd->splitvalue, x);
  else dlo = kd_plane_dist(t, nd->splitdim, nd->splitvalue, x);
  if (dhi <= rad) kd_neighbors_rec(t, nd->right, x, rad, out, cnt, cap);
  if (dlo <= rad) kd_neighbors_rec(t, nd->left, x, rad, out, cnt, cap);
}

/* KDTree.neighbors (kdtree.py:365-372) for ONE query, traversal order; returns the hit count */
ORC_API int64_t orc_kd_neighbors(const orc_kdtree* t, const double* q, double rad, int64_t* out,
                                 int64_t cap) {
  int64_t cnt = 0;
  if (t->root) kd_neighbors_rec(t, t->root, q, rad, out, &cnt, cap);
  return cnt;
}

/* ------------------------------------------------------------------ feasibility --------- */

/* BoxSet.contains (sets.py:177-182) over all coordinates, then the 2-D obstacles of
   Geometric2DCSpace.feasible (geometric.py:56-60): Box = BoxSet.contains, Circle (:15-16). */
static int feasible_pt(const pomp_space_desc* s, const double* x) {
  for (int i = 0; i < s->dim; i++)
    if (x[i] < s->lo[i] || x[i] > s->hi[i]) return 0;
  if (s->n_boxes || s->n_circles) {
    double p[2] = {x[s->obstacle_dim0], x[s->obstacle_dim1]};
    for (int o = 0; o < s->n_boxes; o++) {
      const double* b = s->boxes_h + 4 * o;
      int inside = 1;
      if (p[0] < b[0] || p[0] > b[2]) inside = 0;
      else if (p[1] < b[1] || p[1] > b[3]) inside = 0;
      if (inside) return 0;
    }
    for (int o = 0; o < s->n_circles; o++) {
      const double* c = s->circles_h + 3 * o;
      if (euclid(p, c, 2) <= c[2]) return 0;
    }
  }
  return 1;
}

ORC_API void orc_feasible(const pomp_space_desc* s, const double* x, int64_t n, uint8_t* ok) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) ok[i] = (uint8_t)feasible_pt(s, x + i * s->dim);
}

/* ------------------------------------------------------------------ edge checks --------- */

/* EpsilonEdgeChecker.feasible (edgechecker.py:20-30) with LinearInterpolator
   (interpolators.py:28-33; eval = madd(a, sub(b,a), u): vectorops.py:14-18,115-117) */
static int edge_linear(const pomp_space_desc* s, const double* a, const double* b, double res) {
  int D = s->dim;
  double l = euclid(a, b, D);
  int64_t k = (int64_t)ceil(l / res);
  if (!feasible_pt(s, a) || !feasible_pt(s, b)) return 0;
  double x[POMP_MAX_DIM];
  for (int64_t i = 0; i < k; i++) {
    double u = (double)(i + 1) / (double)(k + 2);
    for (int j = 0; j < D; j++) x[j] = a[j] + u * (b[j] - a[j]);
    if (!feasible_pt(s, x)) return 0;
  }
  return 1;
}

ORC_API void orc_edge_check(const pomp_space_desc* s, const double* a, const double* b, int64_t n,
                            double res, uint8_t* ok) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; i++) ok[i] = (uint8_t)edge_linear(s, a + i * s->dim, b + i * s->dim, res);
}

ORC_API void orc_edge_check_indexed(const pomp_space_desc* s, const double* nodes, const int32_t* ij,
                                    int64_t n, double res, uint8_t* ok) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t e = 0; e < n; e++)
    ok[e] = (uint8_t)edge_linear(s, nodes + (int64_t)ij[2 * e] * s->dim,
                                 nodes + (int64_t)ij[2 * e + 1] * s->dim, res);
}

/* geodesic.interpolate: GeodesicSpace (geodesicspace.py:9-10) / MultiGeodesicSpace (:53-60) /
   SO2Geodesic (so2space.py:11-12 -> so2.interp, so2.py:39-42) */
static void geodesic_interp(const pomp_metric_desc* g, const double* a, const double* b, double u, int D,
                            double* x) {
  if (g->kind == POMP_METRIC_GEODESIC_PRODUCT) {
    int i = 0;
    for (int c = 0; c < g->n_comp; c++) {
      if (g->comp_kind[c] == POMP_COMP_SO2) {
        x[i] = a[i] + so2_diff(b[i], a[i]) * u;
        i++;
      } else {
        for (int j = 0; j < g->comp_dim[c]; j++, i++) x[i] = a[i] + u * (b[i] - a[i]);
      }
    }
    for (; i < D; i++) x[i] = a[i] + u * (b[i] - a[i]);
  } else {
    for (int j = 0; j < D; j++) x[j] = a[j] + u * (b[j] - a[j]);
  }
}

/* EpsilonEdgeChecker.feasible with PiecewiseLinearInterpolator (interpolators.py:73-102):
   length = sum of geodesic segment lengths (:88-92); eval(u) (:93-102) */
static int edge_path(const pomp_space_desc* s, const pomp_metric_desc* g, const double* wp, int np,
                     double res) {
  int D = s->dim;
  double l = 0; /* l = 0 ; l += distance */
  for (int i = 0; i + 1 < np; i++) l += base_metric(g, wp + i * D, wp + (i + 1) * D, D);
  int64_t k = (int64_t)ceil(l / res);
  if (!feasible_pt(s, wp) || !feasible_pt(s, wp + (np - 1) * D)) return 0;
  double x[POMP_MAX_DIM];
  for (int64_t i = 0; i < k; i++) {
    double u = (double)(i + 1) / (double)(k + 2);
    const double* px;
    if (u <= 0) px = wp;
    else if (u >= 1) px = wp + (np - 1) * D;
    else {
      double kk = u * (double)(np - 1);
      double fl = floor(kk);
      double sfrac = kk - fl;
      int seg = (int)fl;
      geodesic_interp(g, wp + seg * D, wp + (seg + 1) * D, sfrac, D, x);
      px = x;
    }
    if (!feasible_pt(s, px)) return 0;
  }
  return 1;
}

ORC_API void orc_edge_check_path(const pomp_space_desc* s, const pomp_metric_desc* g, const double* wp,
                                 const int32_t* off, int64_t n_paths, double res, uint8_t* ok) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t p = 0; p < n_paths; p++)
    ok[p] = (uint8_t)edge_path(s, g, wp + (int64_t)off[p] * s->dim, off[p + 1] - off[p], res);
}

/* ------------------------------------------------------------------ propagation --------- */

static int cmp3(double x, double y) { return x < y ? -1 : (x > y ? 1 : 0); } /* dubins.py:12-15 */

/* DubinsCarSpace.nextState (dubins.py:107-125) */
static void dubins_next(const double* x, const double* u, double* out) {
  double pos[2] = {x[0], x[1]};
  double fwd[2] = {cos(x[2]), sin(x[2])};
  double right[2] = {-fwd[1], fwd[0]};
  double phi = u[1], d = u[0];
  if (fabs(phi) < 1e-8) {
    out[0] = pos[0] + d * fwd[0];
    out[1] = pos[1] + d * fwd[1];
    out[2] = x[2];
    return;
  }
  double c = 1.0 / phi;
  double cor[2] = {pos[0] + c * right[0], pos[1] + c * right[1]};
  int sign = cmp3(d * phi, 0);
  d = fabs(d);
  phi = fabs(phi);
  double thetaMax = d * phi;
  double ang = sign * thetaMax;
  double cs = cos(ang), sn = sin(ang);
  double v[2] = {pos[0] - cor[0], pos[1] - cor[1]};
  /* so2.apply (so2.py:17-21) then vectorops.add = sum([r, cor]) starting from int 0 */
  double r0 = cs * v[0] - sn * v[1], r1 = sn * v[0] + cs * v[1];
  out[0] = (0 + r0) + cor[0];
  out[1] = (0 + r1) + cor[1];
  out[2] = so2_normalize(x[2] + sign * thetaMax);
}

/* ControlSpace.nextState on the base state (no cost coordinate) */
static void next_state_base(const pomp_dyn_desc* dy, const double* x, const double* u, double* out) {
  int X = dy->x_dim;
  switch (dy->kind) {
    case POMP_DYN_ADAPTOR: /* controlspace.py:78-79 */
      for (int i = 0; i < X; i++) out[i] = u[i];
      break;
    case POMP_DYN_CV: { /* statespace.py:62-70 */
      int n = X / 2;
      double dt = u[0];
      const double* a = u + 1;
      double h = 0.5 * pow(dt, 2.0);
      for (int i = 0; i < n; i++) {
        double vnew = x[n + i] + dt * a[i];
        double m1 = x[i] + h * a[i];
        double m2 = x[n + i] * dt;
        out[i] = (0 + m1) + m2; /* vectorops.add -> sum([m1,m2]) */
        out[n + i] = vnew;
      }
      break;
    }
    case POMP_DYN_CV_EULER: { /* statespace.py:47-61 trajectory()[-1] */
      int n = X / 2;
      double duration = u[0];
      const double* a = u + 1;
      double cur[POMP_MAX_DIM], nxt[POMP_MAX_DIM];
      memcpy(cur, x, sizeof(double) * (size_t)X);
      double t = 0.0;
      while (t < duration) {
        double dt = fmin(dy->dt, duration - t);
        double h = 0.5 * pow(dt, 2.0);
        for (int i = 0; i < n; i++) {
          nxt[n + i] = cur[n + i] + dt * a[i];
          double m1 = cur[i] + h * a[i];
          double m2 = cur[n + i] * dt;
          nxt[i] = (0 + m1) + m2;
        }
        memcpy(cur, nxt, sizeof(double) * (size_t)X);
        t = fmin(t + dy->dt, duration);
      }
      memcpy(out, cur, sizeof(double) * (size_t)X);
      break;
    }
    case POMP_DYN_DUBINS:
      dubins_next(x, u, out);
      break;
    case POMP_DYN_PENDULUM: { /* controlspace.py:249-267 with pendulum.py:152-154 */
      double duration = u[0];
      double g = dy->p[0], m = dy->p[1], L = dy->p[2];
      double th = x[0], om = x[1];
      double dt0 = dy->dt;
      if (duration / dy->dt > 10000) dt0 = duration / 10000; /* MAX_INTEGRATION_STEPS, :9,256-258 */
      double t = 0.0;
      while (t < duration) {
        double dth = om;
        double dom = u[1] / (m * pow(L, 2.0)) - g * sin(th) / L;
        double dt = fmin(dt0, duration - t);
        th = th + dt * dth;
        om = om + dt * dom;
        t = fmin(t + dt0, duration);
      }
      out[0] = th;
      out[1] = om;
      break;
    }
    case POMP_DYN_LTI: { /* controlspace.py:101-102: A.dot(x) + B.dot(u) (numpy dot: sequential sum) */
      int U = dy->u_dim;
      for (int i = 0; i < X; i++) {
        double ax = 0, bu = 0;
        for (int j = 0; j < X; j++) ax += dy->A[i * X + j] * x[j];
        for (int j = 0; j < U; j++) bu += dy->B[i * U + j] * u[j];
        out[i] = ax + bu;
      }
      break;
    }
    case POMP_DYN_FLAPPY: { /* flappy.py:20-31 with amount = 1.0 */
      double tc = u[0] * 1.0;
      double net = dy->p[1] + u[1] * dy->p[2];
      out[0] = x[0] + dy->p[0] * tc;
      out[1] = (x[1] + x[2] * tc) + (0.5 * net) * pow(tc, 2.0);
      out[2] = x[2] + net * tc;
      break;
    }
  }
}

/* CostControlSpace.nextState (costspace.py:35-38) + ObjectiveFunction.incremental */
static void next_state_full(const pomp_dyn_desc* dy, const double* x, const double* u, double* out) {
  next_state_base(dy, x, u, out);
  int X = dy->x_dim;
  switch (dy->cost_kind) {
    case POMP_COST_NONE: break;
    case POMP_COST_TIME: out[X] = x[X] + u[0]; break;              /* objectives.py:78-83 */
    case POMP_COST_ABS_U0: out[X] = x[X] + (0 + fabs(u[0])); break; /* dubins.py:131-135: sum(abs(u[0])) */
    case POMP_COST_PATH_LENGTH: out[X] = x[X] + euclid(x, u, X); break; /* objectives.py:25-26 */
  }
}

ORC_API void orc_next_state(const pomp_dyn_desc* dy, const double* x, const double* u, int64_t n,
                            double* xnext) {
  int XF = dy->x_dim + (dy->cost_kind != POMP_COST_NONE);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) next_state_full(dy, x + i * XF, u + i * dy->u_dim, xnext + i * XF);
}

/* RandomControlSelector.select (kinodynamicplanner.py:66-84): strict-< arg-min over the valid
   candidates of metric(nextState(x,u), xdesired) */
ORC_API void orc_select_control(const pomp_dyn_desc* dy, const pomp_metric_desc* m, const double* x,
                                const double* xdes, const double* u, const uint8_t* valid, int64_t B,
                                int S, int32_t* best, double* xnext_best, double* dist_best) {
  int XF = dy->x_dim + (dy->cost_kind != POMP_COST_NONE);
  int U = dy->u_dim;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; b++) {
    double dbest = INFINITY;
    int32_t ib = -1;
    double xn[POMP_MAX_DIM], xb[POMP_MAX_DIM];
    for (int s = 0; s < S; s++) {
      if (valid && !valid[b * S + s]) continue;
      next_state_full(dy, x + b * XF, u + (b * S + s) * U, xn);
      double d = orc_metric(m, xn, xdes + b * XF);
      if (d < dbest) {
        dbest = d;
        ib = s;
        memcpy(xb, xn, sizeof(double) * (size_t)XF);
      }
    }
    best[b] = ib;
    dist_best[b] = dbest;
    for (int i = 0; i < XF; i++) xnext_best[b * XF + i] = ib >= 0 ? xb[i] : NAN;
  }
}

/* Edge check of a control edge, controlSpace.interpolator(x,u) fed to EpsilonEdgeChecker:
   - DUBINS: DubinsCarInterpolator (dubins.py:62-73): length = |u[0]|, eval(s) = nextState(x,[u0*s,u1])
   - PENDULUM / CV_EULER: PiecewiseLinearInterpolator(trajectory(x,u)) with the C-space geodesic
     (controlspace.py:149-172, 249-271; statespace.py:47-61)
   - ADAPTOR: LinearInterpolator(x,u) (controlspace.py:80-81)
   x is the BASE state (no cost coordinate: CostEdgeChecker checks components[0],
   kinodynamicplanner.py:895-896). */
static int edge_control(const pomp_space_desc* s, const pomp_dyn_desc* dy, const pomp_metric_desc* g,
                        const double* x, const double* u, double res) {
  int X = dy->x_dim;
  if (dy->kind == POMP_DYN_ADAPTOR) return edge_linear(s, x, u, res);
  if (dy->kind == POMP_DYN_DUBINS) {
    double xe[3], xs[3];
    dubins_next(x, u, xe);
    double l = fabs(u[0]);
    int64_t k = (int64_t)ceil(l / res);
    if (!feasible_pt(s, x) || !feasible_pt(s, xe)) return 0;
    for (int64_t i = 0; i < k; i++) {
      double uu = (double)(i + 1) / (double)(k + 2);
      double us[2] = {u[0] * uu, u[1]};
      dubins_next(x, us, xs);
      if (!feasible_pt(s, xs)) return 0;
    }
    return 1;
  }
  /* integrated kinds: materialise the trajectory */
  double duration = u[0];
  double dt0 = dy->dt;
  if (dy->kind == POMP_DYN_PENDULUM && duration / dy->dt > 10000) dt0 = duration / 10000;
  int64_t cap = 16, np = 1;
  double* path = (double*)malloc(sizeof(double) * (size_t)(cap * X));
  memcpy(path, x, sizeof(double) * (size_t)X);
  double t = 0.0;
  while (t < duration) {
    if (np == cap) {
      cap *= 2;
      path = (double*)realloc(path, sizeof(double) * (size_t)(cap * X));
    }
    double dt = fmin(dt0, duration - t);
    const double* cur = path + (np - 1) * X;
    double* nxt = path + np * X;
    if (dy->kind == POMP_DYN_PENDULUM) {
      double dth = cur[1];
      double dom = u[1] / (dy->p[1] * pow(dy->p[2], 2.0)) - dy->p[0] * sin(cur[0]) / dy->p[2];
      nxt[0] = cur[0] + dt * dth;
      nxt[1] = cur[1] + dt * dom;
    } else { /* CV_EULER */
      int n = X / 2;
      double h = 0.5 * pow(dt, 2.0);
      for (int i = 0; i < n; i++) {
        nxt[n + i] = cur[n + i] + dt * u[1 + i];
        double m1 = cur[i] + h * u[1 + i];
        double m2 = cur[n + i] * dt;
        nxt[i] = (0 + m1) + m2;
      }
    }
    np++;
    t = fmin(t + dt0, duration);
  }
  int ok = edge_path(s, g, path, (int)np, res);
  free(path);
  return ok;
}

ORC_API void orc_edge_check_control(const pomp_space_desc* s, const pomp_dyn_desc* dy,
                                    const pomp_metric_desc* g, const double* x, const double* u, int64_t n,
                                    double res, uint8_t* ok) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; i++)
    ok[i] = (uint8_t)edge_control(s, dy, g, x + i * dy->x_dim, u + i * dy->u_dim, res);
}

/* ------------------------------------------------------------------ RRT extend ---------- */
/* One kinematic RRT.expand step after sampling (kinodynamicplanner.py:368-411):
   pickNode -> nearest (:451-461), KinematicControlSelector.select (:95-100), LinearInterpolator
   edge (controlspace.py:80-81), EpsilonEdgeChecker (edgechecker.py:20-30).  Does not append. */
ORC_API void orc_rrt_extend(const pomp_metric_desc* m, const double* pts, int64_t n, const uint8_t* mask,
                            const pomp_space_desc* s, const double* xrand, double max_step, double res,
                            int64_t* parent, double* xnew, int32_t* added) {
  int D = m->dim;
  double dist;
  orc_nn_nearest(m, pts, n, mask, xrand, 1, parent, &dist);
  *added = 0;
  if (*parent < 0) return;
  const double* xn = pts + *parent * D;
  double d = euclid(xn, xrand, D); /* cspace.distance */
  if (d > max_step) {
    double u = max_step / d;
    for (int j = 0; j < D; j++) xnew[j] = xn[j] + u * (xrand[j] - xn[j]);
  } else {
    for (int j = 0; j < D; j++) xnew[j] = xrand[j];
  }
  *added = edge_linear(s, xn, xnew, res);
}

/* Thread count of the OpenMP loops (bench.py's CPU arm: torchrun exports OMP_NUM_THREADS=1, which would
 * time the baseline on one core while reporting all of them).  Returns the count now in effect. */
ORC_API int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}
