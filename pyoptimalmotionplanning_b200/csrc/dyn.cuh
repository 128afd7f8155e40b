// dyn.cuh -- ControlSpace.nextState families on the device (one (x,u) pair per thread).
// Operation order follows the reference line by line (see each case); x**2 is evaluated as x*x
// and sin/cos come from the CUDA math library, hence the 1e-5 relative state tolerance of the
// north star instead of bit equality for the trigonometric kinds.
#pragma once
#include "common.cuh"

namespace pomp {

__device__ __forceinline__ int cmp3(double x, double y) { return x < y ? -1 : (x > y ? 1 : 0); }

// DubinsCarSpace.nextState (example_problems/dubins.py:107-125)
__device__ __forceinline__ void dubins_next(const double* x, double d, double phi, double* out) {
  double pos0 = x[0], pos1 = x[1];
  double s0, c0;
  sincos(x[2], &s0, &c0);
  double fwd0 = c0, fwd1 = s0;
  double right0 = -fwd1, right1 = fwd0;
  if (fabs(phi) < 1e-8) {
    out[0] = pos0 + d * fwd0;
    out[1] = pos1 + d * fwd1;
    out[2] = x[2];
    return;
  }
  double c = 1.0 / phi;
  double cor0 = pos0 + c * right0, cor1 = pos1 + c * right1;
  int sign = cmp3(d * phi, 0.0);
  d = fabs(d);
  phi = fabs(phi);
  double thetaMax = d * phi;
  double ang = (double)sign * thetaMax;
  double sn, cs;
  sincos(ang, &sn, &cs);
  double v0 = pos0 - cor0, v1 = pos1 - cor1;
  double r0 = cs * v0 - sn * v1, r1 = sn * v0 + cs * v1;  // so2.apply (klampt/so2.py:17-21)
  out[0] = r0 + cor0;
  out[1] = r1 + cor1;
  out[2] = so2_normalize(x[2] + (double)sign * thetaMax);
}

// one explicit step of the integrated kinds: nxt = step(cur) over dt
__device__ __forceinline__ void euler_step(const pomp_dyn_desc& dy, const double* cur, const double* u, double dt,
                                           double* nxt) {
  if (dy.kind == POMP_DYN_PENDULUM) {  // pendulum.py:152-154 + madd (controlspace.py:263)
    double g = dy.p[0], m = dy.p[1], L = dy.p[2];
    double dth = cur[1];
    double dom = u[1] / (m * (L * L)) - g * sin(cur[0]) / L;
    nxt[0] = cur[0] + dt * dth;
    nxt[1] = cur[1] + dt * dom;
  } else {  // CV_EULER: statespace.py:55-59
    int n = dy.x_dim / 2;
    double hh = 0.5 * (dt * dt);
    for (int i = 0; i < n; i++) {
      double vnew = cur[n + i] + dt * u[1 + i];
      double m1 = cur[i] + hh * u[1 + i];
      double m2 = cur[n + i] * dt;
      nxt[i] = m1 + m2;
      nxt[n + i] = vnew;
    }
  }
}

__device__ __forceinline__ double euler_dt0(const pomp_dyn_desc& dy, double duration) {
  // MAX_INTEGRATION_STEPS guard of LambdaKinodynamicSpace (controlspace.py:9,256-258); CVControlSpace has none
  if (dy.kind == POMP_DYN_PENDULUM && duration / dy.dt > 10000.0) return duration / 10000.0;
  return dy.dt;
}

// base-state transition (no cost coordinate)
__device__ __forceinline__ void next_state_base(const pomp_dyn_desc& dy, const double* x, const double* u,
                                                double* out) {
  const int X = dy.x_dim;
  switch (dy.kind) {
    case POMP_DYN_ADAPTOR:  // controlspace.py:78-79
      for (int i = 0; i < X; i++) out[i] = u[i];
      break;
    case POMP_DYN_CV: {  // statespace.py:62-70
      int n = X / 2;
      double dt = u[0];
      double hh = 0.5 * (dt * dt);
      for (int i = 0; i < n; i++) {
        double vnew = x[n + i] + dt * u[1 + i];
        double m1 = x[i] + hh * u[1 + i];
        double m2 = x[n + i] * dt;
        out[i] = m1 + m2;
        out[n + i] = vnew;
      }
      break;
    }
    case POMP_DYN_CV_EULER:
    case POMP_DYN_PENDULUM: {  // trajectory()[-1]: statespace.py:47-61 / controlspace.py:249-267
      double duration = u[0];
      double dt0 = euler_dt0(dy, duration);
      double cur[POMP_MAX_DIM], nxt[POMP_MAX_DIM];
      for (int i = 0; i < X; i++) cur[i] = x[i];
      double t = 0.0;
      while (t < duration) {
        double dt = fmin(dt0, duration - t);
        euler_step(dy, cur, u, dt, nxt);
        for (int i = 0; i < X; i++) cur[i] = nxt[i];
        t = fmin(t + dt0, duration);
      }
      for (int i = 0; i < X; i++) out[i] = cur[i];
      break;
    }
    case POMP_DYN_DUBINS:
      dubins_next(x, u[0], u[1], out);
      break;
    case POMP_DYN_LTI: {  // controlspace.py:101-102
      int U = dy.u_dim;
      for (int i = 0; i < X; i++) {
        double ax = 0, bu = 0;
        for (int j = 0; j < X; j++) ax += dy.A[i * X + j] * x[j];
        for (int j = 0; j < U; j++) bu += dy.B[i * U + j] * u[j];
        out[i] = ax + bu;
      }
      break;
    }
    case POMP_DYN_FLAPPY: {  // flappy.py:20-31, amount = 1
      double tc = u[0] * 1.0;
      double net = dy.p[1] + u[1] * dy.p[2];
      out[0] = x[0] + dy.p[0] * tc;
      out[1] = (x[1] + x[2] * tc) + (0.5 * net) * (tc * tc);
      out[2] = x[2] + net * tc;
      break;
    }
  }
}

// CostControlSpace.nextState (costspace.py:35-38): append cost + objective.incremental(x,u)
__device__ __forceinline__ void next_state_full(const pomp_dyn_desc& dy, const double* x, const double* u,
                                                double* out) {
  next_state_base(dy, x, u, out);
  const int X = dy.x_dim;
  switch (dy.cost_kind) {
    case POMP_COST_NONE: break;
    case POMP_COST_TIME: out[X] = x[X] + u[0]; break;
    case POMP_COST_ABS_U0: out[X] = x[X] + fabs(u[0]); break;
    case POMP_COST_PATH_LENGTH: {
      double s = 0.0;
      for (int i = 0; i < X; i++) {
        double t = x[i] - u[i];
        s = s + t * t;
      }
      out[X] = x[X] + sqrt(s);
      break;
    }
  }
}

// geodesic.interpolate(a,b,u): geodesicspace.py:9-10,53-60 ; so2space.py:11-12 -> so2.interp
__device__ __forceinline__ void geodesic_interp(const pomp_metric_desc& g, const double* a, const double* b,
                                                double u, int D, double* x) {
  if (g.kind == POMP_METRIC_GEODESIC_PRODUCT) {
    int i = 0;
    for (int c = 0; c < g.n_comp; c++) {
      if (g.comp_kind[c] == POMP_COMP_SO2) {
        x[i] = a[i] + so2_diff(b[i], a[i]) * u;
        i++;
      } else {
        for (int j = 0; j < g.comp_dim[c]; j++, i++) x[i] = a[i] + u * (b[i] - a[i]);
      }
    }
    for (; i < D; i++) x[i] = a[i] + u * (b[i] - a[i]);
  } else {
    for (int j = 0; j < D; j++) x[j] = a[j] + u * (b[j] - a[j]);
  }
}

}  // namespace pomp
