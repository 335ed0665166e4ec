// Synthetic telemetry frames (host).  SURVEY §8(d) config 2 / config 5.
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
#include <cmath>
#include <cstdint>
#include <utility>

#include "pp_internal.h"

namespace {

struct Rng {
  uint64_t key, ctr;
  static uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  Rng(uint64_t seed, uint64_t frame) : key(mix(mix(seed) ^ (frame * 0xD1B54A32D192ED03ull))), ctr(0) {}
  uint64_t next() { return mix(key + (ctr++) * 0x9E3779B97F4A7C15ull); }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double range(double a, double b) { return a + (b - a) * uni(); }
  int below(int n) { return (int)(next() % (uint64_t)n); }
  bool chance(double p) { return uni() < p; }
};

struct Track {
  const double *t;
  int n;
  const double *row(int i) const { return t + (size_t)(((i % n) + n) % n) * PP_MAP_STRIDE; }
  // point on lane `lane` of segment (w-1 -> w) at ratio u, plus tangent / normal
  void at(int w, double u, int lane, double &x, double &y, double &tx, double &ty, double &nx,
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
  void walk(int &w, double &u, int lane, double ds) const {
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

void synth_one(const Track &trk, uint64_t seed, int64_t frame, int n_cars, int rare_permille,
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
  ego_yaw[f] = std::atan2(ty, tx) * 180 / M_PI;
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
      std::swap(cid[a], cid[b]);
      std::swap(cx[a], cx[b]);
      std::swap(cy[a], cy[b]);
      std::swap(cvx[a], cvx[b]);
      std::swap(cvy[a], cvy[b]);
    }
  }
}

}  // namespace

extern "C" int pp_synth_frames(const pp_map *map, uint64_t seed, int64_t first_frame,
                               int64_t n_frames, int32_t n_cars, int32_t rare_permille,
                               const pp_frames *out) {
  if (!map || !out || n_frames < 0) return PP_E_ARG;
  if (n_cars < 0 || n_cars > out->max_cars || out->max_cars > PP_MAX_CARS) return PP_E_RANGE;
  if (!out->ego_x || !out->ego_y || !out->ego_yaw_deg || !out->ego_speed_mph || !out->prev_n ||
      !out->prev_x || !out->prev_y || !out->target_lane_in || !out->n_cars || !out->car_id ||
      !out->car_x || !out->car_y || !out->car_vx || !out->car_vy)
    return PP_E_ARG;
  Track trk{map->table.data(), map->n};
  for (int64_t f = 0; f < n_frames; f++)
    synth_one(trk, seed, first_frame + f, n_cars, rare_permille, out, f);
  return PP_OK;
}
