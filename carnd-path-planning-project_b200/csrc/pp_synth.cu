// Synthetic telemetry frames (host and device twins).  SURVEY §8(d) config 2 / config 5.
//
// The reference has no recorded telemetry, tests or fixtures, so the workload
// the metric is quoted on ("1M independent synthetic frames, 12 cars each,
// 3 lanes") is generated here.  Every frame is a pure function of
// (seed, frame index) through a counter-based generator, so any rank can
// produce any sub-range and the CPU checkers and the GPU path see identical
// inputs.  A configurable share of frames is steered into the branches the
// reference only reaches rarely: standstill / duplicate points,
// off-road, a car > 1000 m away, hard braking, exact ties, cold start,
// car-following (ADJUST/KEEP), collision, crawling ego, ego exactly on a
// waypoint.
//
// The generator is arithmetic only (+ - * /, integer mixing), compiled without FMA
// contraction on both sides, so the device twin (pp_synth_frames_dev: one thread per
// frame, BASELINE config 5 generates its 64M dense frames in HBM instead of shipping
// 217 GB over PCIe) writes the same bits as the host loop.  The one transcendental — the
// yaw, atan2 of a lane segment's direction — is a function of (segment, lane) only and
// comes from a table filled on the host at map creation.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "pp_internal.h"

#define PP_HD __host__ __device__

namespace {

template <class T>
PP_HD inline void swap2(T &a, T &b) {
  const T t = a;
  a = b;
  b = t;
}

struct Rng {
  uint64_t key, ctr;
  PP_HD static uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  PP_HD Rng(uint64_t seed, uint64_t frame) : key(mix(mix(seed) ^ (frame * 0xD1B54A32D192ED03ull))), ctr(0) {}
  PP_HD uint64_t next() { return mix(key + (ctr++) * 0x9E3779B97F4A7C15ull); }
  PP_HD double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  PP_HD double range(double a, double b) { return a + (b - a) * uni(); }
  PP_HD int below(int n) { return (int)(next() % (uint64_t)n); }
  PP_HD bool chance(double p) { return uni() < p; }
};

struct Track {
  const double *t;    // n rows of PP_MAP_STRIDE doubles (un-padded)
  const double *yaw;  // [n][3] heading in degrees of lane segment (w-1 -> w)
  int n;
  PP_HD int wrap(int i) const { return ((i % n) + n) % n; }
  PP_HD const double *row(int i) const { return t + (size_t)(((i % n) + n) % n) * PP_MAP_STRIDE; }
  // point on lane `lane` of segment (w-1 -> w) at ratio u, plus tangent / normal
  PP_HD void at(int w, double u, int lane, double &x, double &y, double &tx, double &ty, double &nx,
          double &ny) const {
    const double *a = row(w - 1), *b = row(w);
    const double ax = a[2 + 2 * lane], ay = a[3 + 2 * lane];
    const double bx = b[2 + 2 * lane], by = b[3 + 2 * lane];
    const double len = b[10 + lane];
    x = ax + (bx - ax) * u;
    y = ay + (by - ay) * u;
    tx = (bx - ax) / len;
    ty = (by - ay) / len;
    nx = b[8];
    ny = b[9];
  }
  // walk ds metres along lane `lane` from (w,u)
  PP_HD void walk(int &w, double &u, int lane, double ds) const {
    for (int guard = 0; guard < 4 * n; guard++) {
      const double len = row(w)[10 + lane];
      if (ds >= 0) {
        const double rem = len * (1 - u);
        if (ds <= rem) {
          u += ds / len;
          return;
        }
        ds -= rem;
        u = 0;
        w++;
      } else {
        const double rem = len * u;
        if (-ds <= rem) {
          u += ds / len;
          return;
        }
        ds += rem;
        u = 1;
        w--;
      }
    }
  }
};

enum Rare {
  R_NONE = -1,
  R_STANDSTILL = 0,
  R_OFFROAD,
  R_FAR_CAR,
  R_HARD_BRAKE,
  R_TIE,
  R_COLD_START,
  R_FOLLOW,
  R_COLLISION,
  R_CRAWL,
  R_ON_WAYPOINT,
  R_EGO_LOST,
  R_COUNT
};

PP_HD void synth_one(const Track &trk, uint64_t seed, int64_t frame, int n_cars, int rare_permille,
               const pp_frames *out, int64_t f) {
  Rng rng(seed, (uint64_t)frame);
  const int mc = out->max_cars;
  double *ego_x = const_cast<double *>(out->ego_x);
  double *ego_y = const_cast<double *>(out->ego_y);
  double *ego_yaw = const_cast<double *>(out->ego_yaw_deg);
  double *ego_mph = const_cast<double *>(out->ego_speed_mph);
  int32_t *prev_n = const_cast<int32_t *>(out->prev_n);
  double *prev_x = const_cast<double *>(out->prev_x) + f * PP_PREV_KEEP;
  double *prev_y = const_cast<double *>(out->prev_y) + f * PP_PREV_KEEP;
  int32_t *tl_in = const_cast<int32_t *>(out->target_lane_in);
  int32_t *ncars = const_cast<int32_t *>(out->n_cars);
  int32_t *cid = const_cast<int32_t *>(out->car_id) + f * mc;
  double *cx = const_cast<double *>(out->car_x) + f * mc;
  double *cy = const_cast<double *>(out->car_y) + f * mc;
  double *cvx = const_cast<double *>(out->car_vx) + f * mc;
  double *cvy = const_cast<double *>(out->car_vy) + f * mc;

  int rare = R_NONE;
  if (rng.below(1000) < rare_permille) rare = rng.below(R_COUNT);

  // ---- ego pose
  int w = rng.below(trk.n);
  double u = rng.uni();
  const int lane = rng.below(PP_NUM_LANES);
  double off = rng.range(-0.5, 0.5);
  double vd = rng.range(-0.3, 0.3);
  if (rng.chance(0.15)) {  // mid lane change
    off = rng.range(-2.0, 2.0);
    vd = rng.range(-2.0, 2.0);
  }
  double v = rng.range(0.5, 22.2);
  double dv = rng.range(-0.1, 0.1);  // per 0.02 s step, i.e. +-5 m/s^2
  if (rare == R_OFFROAD) off = (rng.chance(0.5) ? 1 : -1) * rng.range(21.0, 30.0);
  if (rare == R_HARD_BRAKE) v = rng.range(15.0, 22.2);
  if (rare == R_CRAWL) { v = rng.range(0.02, 0.6); dv = 0; }
  if (rare == R_ON_WAYPOINT) { u = 0; off = 0; vd = 0; }

  double px, py, tx, ty, nx, ny;
  trk.at(w, u, lane, px, py, tx, ty, nx, ny);
  double p9x = px + nx * off, p9y = py + ny * off;
  if (rare == R_ON_WAYPOINT) {  // exactly on a reference waypoint: the tie rule of :172-184
    p9x = trk.row(w - 1)[0];
    p9y = trk.row(w - 1)[1];
  }
  if (rare == R_EGO_LOST) {  // > 1000 m from every lane segment: ego match fails (:1302-1307)
    p9x += 1500.0;
    p9y -= 1300.0;
  }
  double acc_back = 0;
  for (int j = PP_PREV_KEEP - 1; j >= 0; j--) {
    const int back = PP_PREV_KEEP - 1 - j;
    prev_x[j] = p9x - tx * acc_back - nx * (back * vd / 50);
    prev_y[j] = p9y - ty * acc_back - ny * (back * vd / 50);
    double vj = v - back * dv;  // speed of the step arriving at point j
    if (vj < 0.05) vj = 0.05;
    acc_back += vj / 50;
  }
  if (rare == R_STANDSTILL)
    for (int j = 0; j < PP_PREV_KEEP; j++) { prev_x[j] = p9x; prev_y[j] = p9y; }
  prev_n[f] = 47;
  if (rare == R_COLD_START) { const int opts[3] = {0, 3, 9}; prev_n[f] = opts[rng.below(3)]; }
  ego_x[f] = prev_x[0] - tx * (v / 50);
  ego_y[f] = prev_y[0] - ty * (v / 50);
  ego_yaw[f] = trk.yaw[(size_t)trk.wrap(w) * PP_NUM_LANES + lane];  // atan2(ty, tx) in degrees
  ego_mph[f] = v * 2.237;
  int tl = lane;
  if (!rng.chance(0.8)) tl = lane + (rng.chance(0.5) ? 1 : -1);
  if (tl < 0) tl = 0;
  if (tl > PP_NUM_LANES - 1) tl = PP_NUM_LANES - 1;
  tl_in[f] = tl;

  // ---- traffic
  const double ds_lo = n_cars > 16 ? -150.0 : -100.0, ds_hi = n_cars > 16 ? 300.0 : 250.0;
  ncars[f] = n_cars;
  for (int j = 0; j < mc; j++) { cid[j] = 0; cx[j] = cy[j] = cvx[j] = cvy[j] = 0; }
  for (int j = 0; j < n_cars; j++) {
    int cl = rng.below(PP_NUM_LANES);
    double ds = rng.range(ds_lo, ds_hi);
    double jit = rng.range(-0.4, 0.4);
    double sp = (50.0 + rng.range(-10.0, 10.0)) / 2.237;
    if (j == 0) {
      if (rare == R_HARD_BRAKE) { cl = lane; ds = rng.range(6.0, 20.0); sp = rng.range(0.0, 5.0); jit = off; }
      if (rare == R_FOLLOW) { cl = lane; ds = rng.range(12.0, 16.5); sp = v + rng.range(-1.0, 1.0); jit = off; }
      if (rare == R_COLLISION) { cl = lane; ds = rng.range(0.5, 4.0); jit = off; }
      if (rare == R_CRAWL) { cl = lane; ds = rng.range(5.0, 9.0); sp = rng.range(0.0, 0.3); jit = off; }
      if (rare == R_TIE) { cl = lane; ds = rng.range(10.0, 60.0); jit = off; }
    }
    int cw = w;
    double cu = u;
    trk.walk(cw, cu, cl, ds);
    double qx, qy, ux, uy, mx, my;
    trk.at(cw, cu, cl, qx, qy, ux, uy, mx, my);
    cid[j] = j;
    cx[j] = qx + mx * jit;
    cy[j] = qy + my * jit;
    cvx[j] = ux * sp;
    cvy[j] = uy * sp;
  }
  if (rare == R_TIE && n_cars >= 2) {  // same place, same velocity, different id
    cx[1] = cx[0]; cy[1] = cy[0]; cvx[1] = cvx[0]; cvy[1] = cvy[0];
  }
  if (rare == R_FAR_CAR && n_cars >= 1) { cx[0] = p9x + 1500.0; cy[0] = p9y + 900.0; }
  if (rng.chance(0.05)) {  // ids need not arrive in ascending order
    for (int a = 0, b = n_cars - 1; a < b; a++, b--) {
      swap2(cid[a], cid[b]);
      swap2(cx[a], cx[b]);
      swap2(cy[a], cy[b]);
      swap2(cvx[a], cvx[b]);
      swap2(cvy[a], cvy[b]);
    }
  }
}

__global__ void __launch_bounds__(128)
k_synth(Track trk, uint64_t seed, int64_t first_frame, int64_t n_frames, int n_cars,
        int rare_permille, const __grid_constant__ pp_frames out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += stride)
    synth_one(trk, seed, first_frame + f, n_cars, rare_permille, &out, f);
}

int check_out(const pp_map *map, int64_t n_frames, int32_t n_cars, const pp_frames *out) {
  if (!map || !out || n_frames < 0) return PP_E_ARG;
  if (n_cars < 0 || n_cars > out->max_cars || out->max_cars > PP_MAX_CARS) return PP_E_RANGE;
  if (!out->ego_x || !out->ego_y || !out->ego_yaw_deg || !out->ego_speed_mph || !out->prev_n ||
      !out->prev_x || !out->prev_y || !out->target_lane_in || !out->n_cars || !out->car_id ||
      !out->car_x || !out->car_y || !out->car_vx || !out->car_vy)
    return PP_E_ARG;
  return PP_OK;
}

}  // namespace

// Heading (degrees) of every lane segment, the table both generators read.
void ppi::build_yaw_table(const std::vector<double> &table, int n, std::vector<double> &yaw) {
  yaw.assign((size_t)n * PP_NUM_LANES, 0.0);
  for (int w = 0; w < n; w++) {
    const double *a = &table[(size_t)((w - 1 + n) % n) * PP_MAP_STRIDE];
    const double *b = &table[(size_t)w * PP_MAP_STRIDE];
    for (int lane = 0; lane < PP_NUM_LANES; lane++) {
      const double len = b[10 + lane];
      const double tx = (b[2 + 2 * lane] - a[2 + 2 * lane]) / len;
      const double ty = (b[3 + 2 * lane] - a[3 + 2 * lane]) / len;
      yaw[(size_t)w * PP_NUM_LANES + lane] = std::atan2(ty, tx) * 180 / M_PI;
    }
  }
}

extern "C" int pp_synth_frames(const pp_map *map, uint64_t seed, int64_t first_frame,
                               int64_t n_frames, int32_t n_cars, int32_t rare_permille,
                               const pp_frames *out) {
  const int rc = check_out(map, n_frames, n_cars, out);
  if (rc != PP_OK) return rc;
  Track trk{map->table.data(), map->yaw.data(), map->n};
  for (int64_t f = 0; f < n_frames; f++)
    synth_one(trk, seed, first_frame + f, n_cars, rare_permille, out, f);
  return PP_OK;
}

// The same frames written by the GPU into DEVICE buffers (asynchronous on cuda_stream).
extern "C" int pp_synth_frames_dev(const pp_map *map, uint64_t seed, int64_t first_frame,
                                   int64_t n_frames, int32_t n_cars, int32_t rare_permille,
                                   const pp_frames *out_dev, void *cuda_stream) {
  int rc = check_out(map, n_frames, n_cars, out_dev);
  if (rc != PP_OK) return rc;
  if ((rc = ppi::check_map_device(map, "pp_synth_frames_dev")) != PP_OK) return rc;
  if (n_frames == 0) return PP_OK;
  Track trk{map->dev_table + (size_t)PPD_PAD_ROWS * PP_MAP_STRIDE, map->dev_yaw, map->n};
  const int64_t want = (n_frames + 127) / 128;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  k_synth<<<grid, 128, 0, (cudaStream_t)cuda_stream>>>(trk, seed, first_frame, n_frames, n_cars,
                                                      rare_permille, *out_dev);
  ppi::count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ppi::set_cuda_error("k_synth", (int)e, cudaGetErrorString(e));
    return PP_E_CUDA;
  }
  return PP_OK;
}
