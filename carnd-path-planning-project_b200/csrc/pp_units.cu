// pp_units.cu — unit-level entry points: one batch kernel per reference
// function, each a thin wrapper around the SAME __device__ code the fused
// kernel runs (pp_device.cuh).  They exist so that every row of the hot-path
// table can be parity-tested in isolation (tests/test_gpu_units.py) and so
// that the host façade (include/pp.hpp) can offer the reference's function
// names on top of GPU execution.
#include <cuda_runtime.h>

#include "pp_device.cuh"
#include "pp_internal.h"

namespace {

using namespace ppd;

constexpr int kB = 128;

__global__ void k_pt_seg(const double *px, const double *py, const double *ax, const double *ay,
                         const double *bx, const double *by, double *d2, double *rnom,
                         double *rdenom, double *snom, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SegDist s = pt_seg(px[i], py[i], ax[i], ay[i], bx[i], by[i]);
  d2[i] = s.d2;
  rnom[i] = s.rnom;
  rdenom[i] = s.rdenom;
  snom[i] = s.snom;
}

__global__ void k_init_reference(const double *table, int n_wp, const double *x, const double *y,
                                 int32_t *wp, double *ratio, int64_t n) {
  extern __shared__ __align__(16) double s_map[];
  const MapView m = stage_map(s_map, table, n_wp);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RefState rs;
  init_reference(m, x[i], y[i], rs);
  wp[i] = rs.wp;
  ratio[i * 3 + 0] = rs.ratio[0];
  ratio[i * 3 + 1] = rs.ratio[1];
  ratio[i * 3 + 2] = rs.ratio[2];
}

__global__ void k_lane_matching(const double *table, int n_wp, const double *rx, const double *ry,
                                const double *x, const double *y, const double *vx,
                                const double *vy, int32_t *ok, int32_t *lane, int32_t *next_wp,
                                double *s, double *d, double *vs, double *vd, int64_t n) {
  extern __shared__ __align__(16) double s_map[];
  const MapView m = stage_map(s_map, table, n_wp);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RefState rs;
  init_reference(m, rx[i], ry[i], rs);
  const Match mt = lane_match(m, rs, x[i], y[i]);
  double a = 0, b = 0;
  if (mt.ok) project_speed(m, vx[i], vy[i], mt.wp, a, b);
  ok[i] = mt.ok ? 1 : 0;
  lane[i] = mt.ok ? mt.lane : -1;
  next_wp[i] = mt.ok ? mt.wp : 0;
  s[i] = mt.s;
  d[i] = mt.d;
  vs[i] = a;
  vd[i] = b;
}

__global__ void k_lane_pos(const double *table, int n_wp, const double *rx, const double *ry,
                           const double *s, const int32_t *lane, double *ox, double *oy,
                           int32_t *owp, double *odist, int64_t n) {
  extern __shared__ __align__(16) double s_map[];
  const MapView m = stage_map(s_map, table, n_wp);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RefState rs;
  init_reference(m, rx[i], ry[i], rs);
  double qx, qy, dist;
  int wp;
  lane_pos(m, rs, s[i], lane[i], qx, qy, wp, dist);
  ox[i] = qx;
  oy[i] = qy;
  owp[i] = wp;
  odist[i] = dist;
}

__global__ void k_spline(const double *kx, const double *ky, int nk, const double *q, int nq,
                         double *out, int64_t ns) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  Spline sp;
  sp.n = nk;
  for (int k = 0; k < nk; k++) {
    sp.x[k] = kx[i * nk + k];
    sp.y[k] = ky[i * nk + k];
  }
  spline_fit(sp);
  for (int j = 0; j < nq; j++) out[i * nq + j] = spline_eval(sp, q[i * nq + j]);
}

__global__ void k_closest_wp(const double *x, const double *y, const double *mx, const double *my,
                             int nwp, int32_t *out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = closest_waypoint(x[i], y[i], mx, my, nwp);
}
__global__ void k_next_wp(const double *x, const double *y, const double *th, const double *mx,
                          const double *my, int nwp, int32_t *out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = next_waypoint(x[i], y[i], th[i], mx, my, nwp);
}
__global__ void k_get_frenet(const double *x, const double *y, const double *th, const double *mx,
                             const double *my, int nwp, double *os, double *od, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) get_frenet(x[i], y[i], th[i], mx, my, nwp, os[i], od[i]);
}
__global__ void k_get_xy(const double *s, const double *d, const double *ms, const double *mx,
                         const double *my, int nwp, double *ox, double *oy, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) get_xy(s[i], d[i], ms, mx, my, nwp, ox[i], oy[i]);
}

__global__ void k_lane_change(const __grid_constant__ pp_config cfg, const int32_t *car_id,
                              const double *car_s, const double *car_vs, const int32_t *car_lane,
                              int nc, const int32_t *ego_lane, const int32_t *target_lane,
                              const double *ego_s, const double *ego_vs, const double *dt0,
                              int32_t *out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LaneStats ls;
  lane_stats_init(ls, cfg);
  uint32_t flags = 0;
  for (int j = 0; j < nc; j++) {
    const int lane = car_lane[i * nc + j];
    if (lane < 0 || lane > 2) continue;
    lane_stats_add(ls, cfg, car_id[i * nc + j], lane, car_s[i * nc + j], car_vs[i * nc + j],
                   ego_lane[i], target_lane[i], ego_s[i], ego_vs[i], dt0[i], flags);
  }
  out[i] = lane_stats_decide(ls, cfg, ego_lane[i], target_lane[i]);
}

__global__ void k_limit_speed(const __grid_constant__ pp_config cfg, const double *car_vx,
                              const double *car_vy, const double *next_s, const double *ego_s,
                              const double *ego_speed, const double *ego_acc,
                              const int32_t *in_lane, double *ls_speed, double *ls_time,
                              double *sc_speed_o, double *sc_time_o, uint32_t *flags_o, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t flags = 0;
  double ts, tt;
  limit_speed(cfg, car_vx[i], car_vy[i], next_s[i], ego_s[i], ego_speed[i], ego_acc[i],
              in_lane[i] != 0, ts, tt, flags);
  SpeedCtl sc;
  sc_init(sc, cfg, ego_speed[i]);
  sc_limit(sc, ts, tt);
  ls_speed[i] = ts;
  ls_time[i] = tt;
  sc_speed_o[i] = sc.target;
  sc_time_o[i] = sc.time;
  flags_o[i] = flags;
}

// TrajectoryBuilder's control points alone (:638-768), map frame, as the reference logs them.
__global__ void k_control_points(const double *table, int n_wp, const double *pos_x,
                                 const double *pos_y, const int32_t *target_lane,
                                 const double *ego_d, const double *ego_vd, const double *sc_start,
                                 double *out_x, double *out_y, int32_t *out_n, int64_t n) {
  extern __shared__ __align__(16) double s_map[];
  const MapView m = stage_map(s_map, table, n_wp);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RefState rs;
  init_reference(m, pos_x[i], pos_y[i], rs);
  uint32_t flags = 0;
  double cx[6], cy[6];
  const int ncp = traj_control_points(m, rs, pos_x[i], pos_y[i], target_lane[i], ego_d[i], ego_vd[i],
                                      sc_start[i], flags, cx, cy);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    out_x[i * 6 + k] = k < ncp ? cx[k] : __longlong_as_double(0x7ff8000000000000ll);
    out_y[i * 6 + k] = k < ncp ? cy[k] : __longlong_as_double(0x7ff8000000000000ll);
  }
  out_n[i] = ncp;
}

__global__ void k_trajectory(const double *table, int n_wp, const __grid_constant__ pp_config cfg,
                             const int32_t *prev_n, const double *prev_x, const double *prev_y,
                             const double *ego_x, const double *ego_y, const double *yaw,
                             const int32_t *target_lane, const double *ego_d, const double *ego_vd,
                             const double *sc_start, const double *sc_target,
                             const double *sc_time, double *out_x, double *out_y, int32_t *out_n,
                             uint32_t *out_flags, int64_t n) {
  extern __shared__ __align__(16) double s_map[];
  const MapView m = stage_map(s_map, table, n_wp);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int nprev = prev_n[i] >= PP_PREV_KEEP ? PP_PREV_KEEP : 0;
  const double rx = nprev ? prev_x[i * PP_PREV_KEEP + PP_PREV_KEEP - 1] : ego_x[i];
  const double ry = nprev ? prev_y[i * PP_PREV_KEEP + PP_PREV_KEEP - 1] : ego_y[i];
  RefState rs;
  init_reference(m, rx, ry, rs);
  SpeedCtl sc;
  sc.shift = 0;
  sc.start = sc_start[i];
  sc.target = sc_target[i];
  sc.time = sc_time[i];
  uint32_t flags = 0;
  const int np = build_trajectory(m, cfg, rs, prev_x + i * PP_PREV_KEEP, prev_y + i * PP_PREV_KEEP,
                                  nprev, rx, ry, yaw[i], target_lane[i], ego_d[i], ego_vd[i], sc,
                                  out_x + i * PP_PATH_LEN, out_y + i * PP_PATH_LEN, flags);
  for (int k = np; k < PP_PATH_LEN; k++) {
    out_x[i * PP_PATH_LEN + k] = __longlong_as_double(0x7ff8000000000000ll);
    out_y[i * PP_PATH_LEN + k] = __longlong_as_double(0x7ff8000000000000ll);
  }
  out_n[i] = np;
  out_flags[i] = flags;
}

// Self-test of the exact-arithmetic helpers of pp_device.cuh against the generic
// operations they replace: counts[0] div_by != a/b, counts[1] div50 != a/50,
// counts[2] fmod_near != fmod, counts[3] max |atan2_step - atan2| in ulps.
__device__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ double rnd_double(unsigned long long &st, int min_exp, int max_exp) {
  st = mix64(st);
  const unsigned long long mant = st & 0xFFFFFFFFFFFFFull;
  st = mix64(st);
  const int e = min_exp + (int)(st % (unsigned long long)(max_exp - min_exp + 1));
  const unsigned long long sign = (st >> 40) & 1ull;
  const unsigned long long bits = (sign << 63) | ((unsigned long long)(e + 1023) << 52) | mant;
  return __longlong_as_double((long long)bits);
}
__global__ void k_selftest_math(int64_t n, unsigned long long seed, unsigned long long *counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long st = mix64(seed ^ (unsigned long long)i * 0xD1B54A32D192ED03ull);
  // ordinary magnitudes most of the time, extreme ones sometimes (guards must route them)
  const bool wide = (i & 15) == 0;
  double a = rnd_double(st, wide ? -1000 : -40, wide ? 1000 : 40);
  const double b = rnd_double(st, wide ? -1000 : -40, wide ? 1000 : 40);
  if ((i & 63) == 5) a = (i & 64) ? 0.0 : -0.0;  // signed zeros stay on the short path
  if ((i & 3) == 2) {
    // a numerator whose quotient by 50 lies next to a rounding boundary: 50 * (q + ulp/2),
    // moved by up to two ulps either way (the one-correction x/50 has no slack to spare)
    const double q = a;
    const double u = fabs(__longlong_as_double(__double_as_longlong(q) + 1) - q);
    const double hi = q * 50.0, lo = fma(q, 50.0, -hi);
    double m = hi + (lo + (q < 0 ? -25.0 : 25.0) * u);
    const int k = (int)((i >> 2) % 5) - 2;
    m = __longlong_as_double(__double_as_longlong(m) + k);
    a = m;
  }
  const Rcp r = rcp_make(b);
  const double q = div_by(a, r), qe = a / b;
  if (__double_as_longlong(q) != __double_as_longlong(qe)) atomicAdd(&counts[0], 1ull);
  const double h = div50(a), he = a / 50.0;
  if (__double_as_longlong(h) != __double_as_longlong(he)) atomicAdd(&counts[1], 1ull);
  st = mix64(st);
  const double x = -2.0 + 34.0 * ((double)(st >> 11) * (1.0 / 9007199254740992.0));
  const double m = 2 * PPD_PI;
  const double f = fmod_near(x, m), fe = fmod(x, m);
  if (__double_as_longlong(f) != __double_as_longlong(fe)) atomicAdd(&counts[2], 1ull);
  st = mix64(st);
  double dx = 0.001 + 0.5 * ((double)(st >> 11) * (1.0 / 9007199254740992.0));
  if ((i & 7) == 3) dx = -dx;  // all four quadrants
  st = mix64(st);
  const double dy = fabs(dx) * 3.0 * (2.0 * ((double)(st >> 11) * (1.0 / 9007199254740992.0)) - 1.0);
  const double t1 = atan2_step(dy, dx), t2 = atan2(dy, dx);
  long long d = __double_as_longlong(t1) - __double_as_longlong(t2);
  if (d < 0) d = -d;
  if ((t1 < 0) != (t2 < 0) && t1 != t2) d = 1000;  // different signs: flag loudly
  atomicMax(&counts[3], (unsigned long long)d);
  // ---- the emission kernel's lean versions against the guarded helpers, inside the ranges
  // their flag leaves clear (outside, the frame goes to the complete path anyway)
  {
    const Rcp g = rcp_make(b), l = rcp_lean(b);
    if (g.ok != l.ok || (g.ok && __double_as_longlong(g.y) != __double_as_longlong(l.y)))
      atomicAdd(&counts[4], 1ull);
    if (mag_ok(a)) {
      const double q = div50_raw(a), qe = a / 50.0;
      // a -0 numerator comes back as +0 (documented): compare values, not zero signs
      if (!(q == qe) || (qe != 0 && __double_as_longlong(q) != __double_as_longlong(qe)))
        atomicAdd(&counts[5], 1ull);
    }
    const double l1 = atan2_lean(dy, dx);
    if (__double_as_longlong(l1) != __double_as_longlong(t1)) atomicAdd(&counts[6], 1ull);
    st = mix64(st);
    const double xw = 25.132741228718345 * ((double)(st >> 11) * (1.0 / 9007199254740992.0));  // [0, 8 pi)
    const double w1 = wrap_lean(xw), w2 = fmod(xw, m);
    if (__double_as_longlong(w1) != __double_as_longlong(w2)) atomicAdd(&counts[7], 1ull);
  }
}

__global__ void k_project_speed(const double *table, int n_wp, const double *vx, const double *vy,
                                const int32_t *next_wp, double *ovs, double *ovd, int64_t n) {
  extern __shared__ __align__(16) double s_map[];
  const MapView m = stage_map(s_map, table, n_wp);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a, b;
  project_speed(m, vx[i], vy[i], next_wp[i], a, b);
  ovs[i] = a;
  ovd[i] = b;
}

__global__ void k_speed_controller(int op, const __grid_constant__ pp_config cfg, const double *start,
                                   double *target, double *time, double *shift, const double *a,
                                   const double *b, double *out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SpeedCtl c;
  c.start = start[i];
  c.target = target[i];
  c.time = time[i];
  c.shift = shift[i];
  if (op == 3) {
    sc_init(c, cfg, start[i]);
    target[i] = c.target;
    time[i] = c.time;
    shift[i] = c.shift;
  } else if (op == 0) {
    out[i] = sc_speed(c, a[i]);
  } else if (op == 1) {
    sc_limit(c, a[i], b[i]);
    target[i] = c.target;
    time[i] = c.time;
  } else {
    sc_override(c, a[i], b[i]);
    shift[i] = c.shift;
  }
}

int finish(const char *what) {
  ppi::count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ppi::set_cuda_error(what, (int)e, cudaGetErrorString(e));
    return PP_E_CUDA;
  }
  return PP_OK;
}
inline int grid_for(int64_t n) { return (int)((n + kB - 1) / kB); }
inline size_t map_smem(const pp_map *m) { return map_stage_bytes(m->n); }
int need_map(const pp_map *m, const char *who) {
  const int rc = ppi::check_map_device(m, who);
  if (rc != PP_OK) return rc;
  if (map_smem(m) > 48 * 1024) return PP_E_RANGE;
  return PP_OK;
}

}  // namespace

extern "C" {

int pp_distancesq_pt_seg_batch(const double *px, const double *py, const double *ax,
                               const double *ay, const double *bx, const double *by,
                               double *out_d2, double *out_rnom, double *out_rdenom,
                               double *out_snom, int64_t n, void *stream) {
  if (!px || !py || !ax || !ay || !bx || !by || !out_d2 || !out_rnom || !out_rdenom || !out_snom ||
      n < 0)
    return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_pt_seg<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(px, py, ax, ay, bx, by, out_d2, out_rnom,
                                                         out_rdenom, out_snom, n);
  return finish("k_pt_seg");
}

int pp_init_reference_waypoint_batch(const pp_map *map, const double *x, const double *y,
                                     int32_t *out_ref_wp, double *out_ratio, int64_t n,
                                     void *stream) {
  int rc = need_map(map, "pp_init_reference_waypoint_batch");
  if (rc != PP_OK) return rc;
  if (!x || !y || !out_ref_wp || !out_ratio || n < 0) return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_init_reference<<<grid_for(n), kB, map_smem(map), (cudaStream_t)stream>>>(
      map->dev_table, map->n, x, y, out_ref_wp, out_ratio, n);
  return finish("k_init_reference");
}

int pp_lane_matching_batch(const pp_map *map, const double *rx, const double *ry,
                           const double *x, const double *y, const double *vx, const double *vy,
                           int32_t *out_ok, int32_t *out_lane, int32_t *out_next_wp,
                           double *out_s, double *out_d, double *out_vs, double *out_vd, int64_t n,
                           void *stream) {
  int rc = need_map(map, "pp_lane_matching_batch");
  if (rc != PP_OK) return rc;
  if (!rx || !ry || !x || !y || !vx || !vy || !out_ok || !out_lane || !out_next_wp || !out_s ||
      !out_d || !out_vs || !out_vd || n < 0)
    return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_lane_matching<<<grid_for(n), kB, map_smem(map), (cudaStream_t)stream>>>(
      map->dev_table, map->n, rx, ry, x, y, vx, vy, out_ok, out_lane, out_next_wp, out_s, out_d,
      out_vs, out_vd, n);
  return finish("k_lane_matching");
}

int pp_get_lane_pos_batch(const pp_map *map, const double *rx, const double *ry, const double *s,
                          const int32_t *lane, double *out_x, double *out_y, int32_t *out_wp,
                          double *out_dist, int64_t n, void *stream) {
  int rc = need_map(map, "pp_get_lane_pos_batch");
  if (rc != PP_OK) return rc;
  if (!rx || !ry || !s || !lane || !out_x || !out_y || !out_wp || !out_dist || n < 0)
    return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_lane_pos<<<grid_for(n), kB, map_smem(map), (cudaStream_t)stream>>>(
      map->dev_table, map->n, rx, ry, s, lane, out_x, out_y, out_wp, out_dist, n);
  return finish("k_lane_pos");
}

int pp_spline_batch(const double *kx, const double *ky, int32_t n_knots, const double *q,
                    int32_t n_q, double *out, int64_t n_splines, void *stream) {
  if (!kx || !ky || !q || !out || n_splines < 0 || n_q < 0) return PP_E_ARG;
  if (n_knots < 3 || n_knots > 15) return PP_E_RANGE;
  if (n_splines == 0) return PP_OK;
  k_spline<<<grid_for(n_splines), kB, 0, (cudaStream_t)stream>>>(kx, ky, n_knots, q, n_q, out,
                                                                 n_splines);
  return finish("k_spline");
}

int pp_lane_change_batch(const pp_config *cfg, const int32_t *car_id, const double *car_s,
                         const double *car_vs, const int32_t *car_lane, int32_t n_cars,
                         const int32_t *ego_lane, const int32_t *target_lane, const double *ego_s,
                         const double *ego_vs, const double *dt0, int32_t *out_target_lane,
                         int64_t n, void *stream) {
  if (!cfg || !ego_lane || !target_lane || !ego_s || !ego_vs || !dt0 || !out_target_lane || n < 0 ||
      n_cars < 0)
    return PP_E_ARG;
  if (n_cars > 0 && (!car_id || !car_s || !car_vs || !car_lane)) return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_lane_change<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(*cfg, car_id, car_s, car_vs, car_lane,
                                                              n_cars, ego_lane, target_lane, ego_s,
                                                              ego_vs, dt0, out_target_lane, n);
  return finish("k_lane_change");
}

int pp_limit_speed_batch(const pp_config *cfg, const double *car_vx, const double *car_vy,
                         const double *next_s, const double *ego_s, const double *ego_speed,
                         const double *ego_acc, const int32_t *in_lane, double *out_ls_speed,
                         double *out_ls_time, double *out_sc_speed, double *out_sc_time,
                         uint32_t *out_flags, int64_t n, void *stream) {
  if (!cfg || !car_vx || !car_vy || !next_s || !ego_s || !ego_speed || !ego_acc || !in_lane ||
      !out_ls_speed || !out_ls_time || !out_sc_speed || !out_sc_time || !out_flags || n < 0)
    return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_limit_speed<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(
      *cfg, car_vx, car_vy, next_s, ego_s, ego_speed, ego_acc, in_lane, out_ls_speed, out_ls_time,
      out_sc_speed, out_sc_time, out_flags, n);
  return finish("k_limit_speed");
}

int pp_trajectory_build_batch(const pp_map *map, const pp_config *cfg, const int32_t *prev_n,
                              const double *prev_x, const double *prev_y, const double *ego_x,
                              const double *ego_y, const double *ego_yaw_deg,
                              const int32_t *target_lane, const double *ego_d,
                              const double *ego_vd, const double *sc_start,
                              const double *sc_target, const double *sc_time, double *out_x,
                              double *out_y, int32_t *out_n, uint32_t *out_flags, int64_t n,
                              void *stream) {
  int rc = need_map(map, "pp_trajectory_build_batch");
  if (rc != PP_OK) return rc;
  if (!cfg || !prev_n || !prev_x || !prev_y || !ego_x || !ego_y || !ego_yaw_deg || !target_lane ||
      !ego_d || !ego_vd || !sc_start || !sc_target || !sc_time || !out_x || !out_y || !out_n ||
      !out_flags || n < 0)
    return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_trajectory<<<grid_for(n), kB, map_smem(map), (cudaStream_t)stream>>>(
      map->dev_table, map->n, *cfg, prev_n, prev_x, prev_y, ego_x, ego_y, ego_yaw_deg, target_lane,
      ego_d, ego_vd, sc_start, sc_target, sc_time, out_x, out_y, out_n, out_flags, n);
  return finish("k_trajectory");
}

int pp_control_points_batch(const pp_map *map, const double *pos_x, const double *pos_y,
                            const int32_t *target_lane, const double *ego_d, const double *ego_vd,
                            const double *sc_start, double *out_x, double *out_y, int32_t *out_n,
                            int64_t n, void *stream) {
  int rc = need_map(map, "pp_control_points_batch");
  if (rc != PP_OK) return rc;
  if (!pos_x || !pos_y || !target_lane || !ego_d || !ego_vd || !sc_start || !out_x || !out_y ||
      !out_n || n < 0)
    return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_control_points<<<grid_for(n), kB, map_smem(map), (cudaStream_t)stream>>>(
      map->dev_table, map->n, pos_x, pos_y, target_lane, ego_d, ego_vd, sc_start, out_x, out_y,
      out_n, n);
  return finish("k_control_points");
}

int pp_project_speed_batch(const pp_map *map, const double *vx, const double *vy,
                           const int32_t *next_wp, double *out_vs, double *out_vd, int64_t n,
                           void *stream) {
  if (!vx || !vy || !next_wp || !out_vs || !out_vd || n < 0) return PP_E_ARG;
  int rc = need_map(map, "pp_project_speed_batch");
  if (rc != PP_OK) return rc;
  if (n == 0) return PP_OK;
  k_project_speed<<<grid_for(n), kB, map_smem(map), (cudaStream_t)stream>>>(
      map->dev_table, map->n, vx, vy, next_wp, out_vs, out_vd, n);
  return finish("k_project_speed");
}

int pp_speed_controller_batch(int32_t op, const double *start, double *target, double *time,
                              double *shift, const double *a, const double *b, double *out,
                              int64_t n, void *stream) {
  if (op < 0 || op > 3 || !start || !target || !time || !shift || n < 0) return PP_E_ARG;
  if (op != 3 && !a) return PP_E_ARG;
  if (op == 0 && !out) return PP_E_ARG;
  if ((op == 1 || op == 2) && !b) return PP_E_ARG;
  if (n == 0) return PP_OK;
  pp_config cfg;
  pp_config_default(&cfg);
  k_speed_controller<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(op, cfg, start, target, time,
                                                                   shift, a, b, out, n);
  return finish("k_speed_controller");
}

int pp_selftest_math(int64_t n, uint64_t seed, int64_t *counts_dev, void *stream) {
  if (!counts_dev || n < 0) return PP_E_ARG;
  if (cudaMemsetAsync(counts_dev, 0, 8 * sizeof(int64_t), (cudaStream_t)stream) != cudaSuccess)
    return finish("pp_selftest_math memset");
  if (n == 0) return PP_OK;
  k_selftest_math<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(
      n, (unsigned long long)seed, (unsigned long long *)counts_dev);
  return finish("k_selftest_math");
}

int pp_closest_waypoint_batch(const double *x, const double *y, const double *maps_x,
                              const double *maps_y, int32_t n_wp, int32_t *out, int64_t n,
                              void *stream) {
  if (!x || !y || !maps_x || !maps_y || !out || n < 0 || n_wp < 1) return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_closest_wp<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(x, y, maps_x, maps_y, n_wp, out, n);
  return finish("k_closest_wp");
}

int pp_next_waypoint_batch(const double *x, const double *y, const double *theta,
                           const double *maps_x, const double *maps_y, int32_t n_wp, int32_t *out,
                           int64_t n, void *stream) {
  if (!x || !y || !theta || !maps_x || !maps_y || !out || n < 0 || n_wp < 1) return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_next_wp<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(x, y, theta, maps_x, maps_y, n_wp, out,
                                                          n);
  return finish("k_next_wp");
}

int pp_get_frenet_batch(const double *x, const double *y, const double *theta,
                        const double *maps_x, const double *maps_y, int32_t n_wp, double *out_s,
                        double *out_d, int64_t n, void *stream) {
  if (!x || !y || !theta || !maps_x || !maps_y || !out_s || !out_d || n < 0 || n_wp < 2)
    return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_get_frenet<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(x, y, theta, maps_x, maps_y, n_wp,
                                                             out_s, out_d, n);
  return finish("k_get_frenet");
}

int pp_get_xy_batch(const double *s, const double *d, const double *maps_s, const double *maps_x,
                    const double *maps_y, int32_t n_wp, double *out_x, double *out_y, int64_t n,
                    void *stream) {
  if (!s || !d || !maps_s || !maps_x || !maps_y || !out_x || !out_y || n < 0 || n_wp < 2)
    return PP_E_ARG;
  if (n == 0) return PP_OK;
  k_get_xy<<<grid_for(n), kB, 0, (cudaStream_t)stream>>>(s, d, maps_s, maps_x, maps_y, n_wp, out_x,
                                                         out_y, n);
  return finish("k_get_xy");
}

}  // extern "C"
