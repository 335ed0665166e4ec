// TEST INFRASTRUCTURE (oracle/_ref).  NOT part of the product; nothing under
// carnd-path-planning-project_b200/ may include, link or call this file.
//
// Socket-free CPU harness around the reference's OWN planner code.  The
// reference translation unit /root/reference/src/main.cpp is compiled
// unmodified (it is #included below from where it lies; no reference source is
// copied into this repo) against the transport stub oracle/stub/uWS/uWS.h.
//   * class harness  : restates the per-frame glue of main::onMessage
//                      (src/main.cpp:1254-1457) around the reference classes
//                      Map / Car / LaneChangePlanner / LimitSpeed /
//                      SpeedController / TrajectoryBuilder / tk::spline and
//                      runs it over the same SoA frame buffers pp_plan_batch
//                      consumes (include/pp.h).
//   * lambda harness : feeds 42["telemetry",{...}] strings through the
//                      UNTOUCHED onMessage lambda (src/main.cpp:1214-1474) and
//                      parses the 42["control",...] reply; used to check the
//                      glue restatement (agreement limited to the 15
//                      significant digits of src/json.hpp:6689-6692).
//   * unit exports   : one C function per reference function, to pin the C
//                      restatement in oracle/pp_oracle.c function by function.
// Built by oracle/Makefile into oracle/_ref/libppref.so with
//   g++ -std=c++11 -O2 -ffp-contract=off   (bit-identical across
//   -O0/-O2; FMA contraction changes results, so it is off).

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../include/pp.h"
#include <uWS/uWS.h>

// ---- print interception: the reference reports anomalies only by printing.
// printf/fprintf inside the reference TU are routed here and turned into the
// PP_F_* bits of include/pp.h.  Nothing the reference computes depends on
// these calls.
static thread_local uint32_t t_flags = 0;

static void classify(const char *fmt, va_list ap) {
  if (!fmt) return;
  const char *p = fmt;
  while (*p == '\n' || *p == ' ') p++;
  if (!strncmp(p, "lane_switch_time(", 17)) t_flags |= PP_F_LANE_SWITCH_NEG;
  else if (!strncmp(p, "spline input error", 18)) t_flags |= PP_F_SPLINE_INPUT_ERR;
  else if (!strncmp(p, "spline warning", 14)) t_flags |= PP_F_SPLINE_WARNING;
  else if (!strncmp(p, "detected collision", 18)) t_flags |= PP_F_COLLISION;
  else if (!strncmp(p, "Warning! can't lane match ego", 29)) t_flags |= PP_F_EGO_MATCH_FAIL;
  else if (!strncmp(p, "Warning! can't lane match car", 29)) t_flags |= PP_F_CAR_DROPPED;
  else if (!strncmp(p, "maxbrake:", 9)) t_flags |= PP_F_MAXBRAKE;
  else if (!strncmp(p, "normalbrake:", 12)) t_flags |= PP_F_BRAKE;
  else if (!strncmp(p, "accT too high", 13)) t_flags |= PP_F_ACCT_HIGH;
  else if (!strncmp(p, "accN too high", 13)) t_flags |= PP_F_ACCN_HIGH;
  else if (!strncmp(p, "acceleration override", 21)) t_flags |= PP_F_ACC_OVERRIDE;
  else if (!strncmp(p, "adjusting curvature", 19)) t_flags |= PP_F_CURV_ADJUST;
  else if (!strncmp(p, "transform calculation error", 27)) t_flags |= PP_F_TRANSFORM_ERR;
  else if (!strncmp(p, "target lane too far", 19)) t_flags |= PP_F_VETO;
  else if (!strncmp(p, "limitspeed %s", 13)) {
    const char *code = va_arg(ap, const char *);
    if (code && !strcmp(code, "ADJUST")) t_flags |= PP_F_ADJUST;
    if (code && !strcmp(code, "KEEP")) t_flags |= PP_F_KEEP;
  }
}
static int ppref_printf(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  classify(fmt, ap);
  va_end(ap);
  return 0;
}
static FILE *g_tee = NULL;  // ppref_plan_frames_log: the reference's log text goes here as well
static int ppref_fprintf(FILE *, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  if (g_tee) {
    va_list ap2;
    va_copy(ap2, ap);
    vfprintf(g_tee, fmt, ap2);
    va_end(ap2);
  }
  classify(fmt, ap);
  va_end(ap);
  return 0;
}

namespace uWS {
SendSink &send_sink() {
  static thread_local SendSink s;
  return s;
}
}  // namespace uWS

#define printf ppref_printf
#define fprintf ppref_fprintf
#define main ref_main
#include "main.cpp"  // = /root/reference/src/main.cpp via -I (unmodified)
#undef main
#undef fprintf
#undef printf

// Bits this harness can observe (the reference prints them); the others
// (FALLBACK, CLOSED_*, COLD_START) are silent in the reference.
static const uint32_t kObservable =
    PP_F_EGO_MATCH_FAIL | PP_F_CAR_DROPPED | PP_F_COLLISION | PP_F_BRAKE | PP_F_MAXBRAKE |
    PP_F_ADJUST | PP_F_KEEP | PP_F_SPLINE_INPUT_ERR | PP_F_ACC_OVERRIDE | PP_F_CURV_ADJUST |
    PP_F_LANE_SWITCH_NEG | PP_F_VETO | PP_F_ACCT_HIGH | PP_F_ACCN_HIGH | PP_F_SPLINE_WARNING |
    PP_F_TRANSFORM_ERR;

struct ppref_map {
  Map map;
};

// ---------------------------------------------------------------------------
// class harness: one frame through the reference classes.
// Glue follows src/main.cpp:1254-1457 statement by statement; the persistent
// std::map<int,Car> is fresh per frame (equivalent when every frame reports
// its full id set) and target_lane comes in / goes out as data.
// ---------------------------------------------------------------------------
static void plan_one(Map &map, const pp_frames *in, const pp_plans *out, int64_t f) {
  const int keep = PP_PREV_KEEP;
  double ex = in->ego_x[f], ey = in->ego_y[f];
  double eyaw = in->ego_yaw_deg[f];
  double espeed = in->ego_speed_mph[f];
  espeed /= 2.237;  // :1239
  double eacc = 0;
  int target_lane = in->target_lane_in[f];
  uint32_t extra_flags = 0;

  vector<Point> prev;
  double dt0 = 0;
  Point esv;  // ego speed vector, (0,0) unless a previous path exists
  if (in->prev_n[f] >= keep) {  // :1261
    for (int i = 0; i < keep; i++)
      prev.push_back(Point(in->prev_x[f * keep + i], in->prev_y[f * keep + i]));
    double v2 = (prev[keep - 2] - prev[keep - 3]).length();
    esv = prev[keep - 1] - prev[keep - 2];
    double v3 = esv.length();
    eacc = (v3 - v2) * 50;
    espeed = v3 * 50;
    esv.x *= 50;
    esv.y *= 50;
    ex = prev[keep - 1].x;
    ey = prev[keep - 1].y;
    dt0 = keep / 50.0;
  } else {
    extra_flags |= PP_F_COLD_START;
  }

  map.init_reference_waypoint(ex, ey);  // :1299
  int elane;
  double es, ed;
  if (!map.lane_matching(ex, ey, es, ed, elane)) {
    ppref_printf("Warning! can't lane match ego\n");
    es = ed = 0;
    elane = 0;
  }
  double evs, evd;
  map.project_speed(esv, map.reference_waypoint_id, &evs, &evd);  // :1313
  if (eacc > maximum_acc) eacc = maximum_acc;  // :1319-1320
  if (eacc < -maximum_acc) eacc = -maximum_acc;

  std::map<int, Car> cars;
  const int mc = in->max_cars;
  const int nc = in->n_cars[f];
  for (int j = 0; j < nc; j++) {  // :1325-1350
    int id = in->car_id[f * mc + j];
    Car &car = cars[id];
    car.id = id;
    car.x = in->car_x[f * mc + j];
    car.y = in->car_y[f * mc + j];
    car.vx = in->car_vx[f * mc + j];
    car.vy = in->car_vy[f * mc + j];
    int nwp = 0;
    bool ok = map.lane_matching(car.x, car.y, car.s, car.d, car.lane, &nwp);
    if (ok) map.project_speed(Point(car.vx, car.vy), nwp, &car.vs, &car.vd);
    if (out->car_lane) out->car_lane[f * mc + j] = ok ? car.lane : -1;
    if (out->car_next_wp) out->car_next_wp[f * mc + j] = ok ? nwp : 0;
    if (out->car_s) out->car_s[f * mc + j] = ok ? car.s : 0;
    if (out->car_d) out->car_d[f * mc + j] = ok ? car.d : 0;
    if (out->car_vs) out->car_vs[f * mc + j] = ok ? car.vs : 0;
    if (out->car_vd) out->car_vd[f * mc + j] = ok ? car.vd : 0;
    if (!ok) {
      ppref_printf("Warning! can't lane match car %d at %.2f %.2f\n", id, car.x, car.y);
      cars.erase(cars.find(id));
    }
  }

  LaneChangePlanner lcp;  // :1352-1356
  target_lane = lcp.calculate_target_lane(cars, elane, target_lane, es, evs, dt0);
  if (target_lane != elane) {  // :1358-1369
    double dtl = map.get_lane_center_offset(target_lane);
    double diff = fabs(evd * 1.0 + ed - dtl);
    if (diff > 6.0) {
      ppref_printf("target lane too far\n");
      target_lane = elane;
    }
  }

  int next_id = -1, next_id_tl = -1;  // :1383-1411
  double next_s = 0, next_s_tl = 0;
  double dtl = map.get_lane_center_offset(target_lane);
  for (auto &kv : cars) {
    Car &o = kv.second;
    double s0 = o.predicted_s(dt0);
    double d0 = o.predicted_d(dt0);
    if (s0 > es && fabs(d0 - ed) < 3) {
      if (next_id == -1 || next_s > s0) {
        next_id = o.id;
        next_s = s0;
      }
    }
    if (s0 >= es - car_length - safety_distance && fabs(d0 - dtl) < 3) {
      if (next_id_tl == -1 || next_s_tl > s0) {
        next_id_tl = o.id;
        next_s_tl = s0;
      }
    }
  }
  if (next_id_tl == next_id) next_id_tl = -1;

  SpeedController sc(espeed);  // :1422-1438
  if (next_id != -1) {
    LimitSpeed ls;
    ls.calculate(cars[next_id], next_s, es, espeed, eacc, true);
    sc.add_limit_breakpoint(ls.target_speed, ls.target_time);
  }
  if (next_id_tl != -1) {
    LimitSpeed ls;
    ls.calculate(cars[next_id_tl], next_s_tl, es, espeed, eacc, false);
    sc.add_limit_breakpoint(ls.target_speed, ls.target_time);
  }
  if (out->target_speed) out->target_speed[f] = sc.target_speed;
  if (out->target_time) out->target_time[f] = sc.target_time;

  TrajectoryBuilder tb;  // :1446-1448
  vector<Point> traj = tb.build(prev, ex, ey, eyaw, elane, target_lane, ed, evd, map, sc);

  const int np = (int)traj.size();
  for (int i = 0; i < np && i < PP_PATH_LEN; i++) {
    out->next_x[f * PP_PATH_LEN + i] = traj[i].x;
    out->next_y[f * PP_PATH_LEN + i] = traj[i].y;
  }
  out->n_points[f] = np;
  out->ego_lane[f] = elane;
  out->ref_wp[f] = map.reference_waypoint_id;
  out->target_lane[f] = target_lane;
  if (out->ego_s) out->ego_s[f] = es;
  if (out->ego_d) out->ego_d[f] = ed;
  if (out->ego_vs) out->ego_vs[f] = evs;
  if (out->ego_vd) out->ego_vd[f] = evd;
  if (out->ego_speed) out->ego_speed[f] = espeed;
  if (out->ego_acc) out->ego_acc[f] = eacc;
  if (out->next_car_id) out->next_car_id[f] = next_id;
  if (out->next_car_in_target_lane) out->next_car_in_target_lane[f] = next_id_tl;
  t_flags |= extra_flags;
}

static FILE *devnull() {
  static FILE *f = fopen("/dev/null", "w");
  return f;
}

extern "C" {

uint32_t ppref_observable_flags(void) { return kObservable; }

ppref_map *ppref_map_create(const double *wx, const double *wy, int n) {
  ppref_map *m = new ppref_map();
  vector<double> xs(wx, wx + n), ys(wy, wy + n);
  m->map.Init(xs, ys);
  return m;
}

// CSV parse as src/main.cpp:1171-1191 (x,y double; s,dx,dy float, ignored).
ppref_map *ppref_map_create_from_csv(const char *path) {
  std::ifstream is(path, std::ifstream::in);
  if (!is.good()) return NULL;
  vector<double> xs, ys;
  string line;
  while (getline(is, line)) {
    std::istringstream iss(line);
    double x, y;
    float s, dx, dy;
    iss >> x;
    iss >> y;
    iss >> s;
    iss >> dx;
    iss >> dy;
    xs.push_back(x);
    ys.push_back(y);
  }
  ppref_map *m = new ppref_map();
  m->map.Init(xs, ys);
  return m;
}

void ppref_map_destroy(ppref_map *m) { delete m; }
int ppref_map_num_waypoints(ppref_map *m) { return (int)m->map.waypoints.size(); }

// Table in the PP_MAP_STRIDE layout of include/pp.h.
void ppref_map_table(ppref_map *m, double *out) {
  Map &map = m->map;
  int n = (int)map.waypoints.size();
  for (int i = 0; i < n; i++) {
    Map::Waypoint &w = map.waypoints[i];
    double *o = out + (size_t)i * PP_MAP_STRIDE;
    o[0] = w.ref.x;
    o[1] = w.ref.y;
    for (int l = 0; l < 3; l++) {
      o[2 + 2 * l] = w.lane_center[l].x;
      o[3 + 2 * l] = w.lane_center[l].y;
    }
    o[8] = w.nx;
    o[9] = w.ny;
    for (int l = 0; l < 3; l++) o[10 + l] = map.get_lane_length(i, l);
  }
}

// Plan n frames with `threads` host threads (each with a private Map copy:
// the reference Map carries per-frame mutable state, src/main.cpp:132-133).
// want_flags != 0 turns the reference's fLog sites on so that every
// observable PP_F_* bit is reported (slower; never used for timing).
int ppref_plan_frames(ppref_map *m, const pp_frames *in, const pp_plans *out, int64_t n,
                      int threads, int want_flags) {
  if (!m || !in || !out) return PP_E_ARG;
  fLog = want_flags ? devnull() : NULL;
  if (threads < 1) threads = 1;
  auto work = [&](int64_t lo, int64_t hi) {
    Map local = m->map;
    for (int64_t f = lo; f < hi; f++) {
      t_flags = 0;
      plan_one(local, in, out, f);
      if (out->flags) out->flags[f] = t_flags;
    }
  };
  if (threads == 1) {
    work(0, n);
  } else {
    std::vector<std::thread> pool;
    int64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
      int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
      if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    for (auto &th : pool) th.join();
  }
  fLog = NULL;
  return PP_OK;
}

// The same with the reference's trajectory.log sites writing to `log_path` (single thread): the
// text the reference's classes print for these frames (control_points=, result=, ...).
int ppref_plan_frames_log(ppref_map *m, const pp_frames *in, const pp_plans *out, int64_t n,
                          const char *log_path) {
  if (!m || !in || !out || !log_path) return PP_E_ARG;
  FILE *f = fopen(log_path, "wt");
  if (!f) return PP_E_IO;
  fLog = f;
  g_tee = f;
  Map local = m->map;
  for (int64_t i = 0; i < n; i++) {
    t_flags = 0;
    fprintf(f, "frame %lld\n", (long long)i);
    plan_one(local, in, out, i);
    if (out->flags) out->flags[i] = t_flags;
  }
  fLog = NULL;
  g_tee = NULL;
  fclose(f);
  return PP_OK;
}

// ---- unit exports --------------------------------------------------------

void ppref_distancesq_pt_seg(const double *px, const double *py, const double *ax,
                             const double *ay, const double *bx, const double *by, double *d2,
                             double *rnom, double *rdenom, double *snom, int64_t n) {
  for (int64_t i = 0; i < n; i++)
    d2[i] = distancesq_pt_seg(Point(px[i], py[i]), Point(ax[i], ay[i]), Point(bx[i], by[i]),
                              rnom[i], rdenom[i], snom[i]);
}

void ppref_init_reference_waypoint(ppref_map *m, const double *x, const double *y,
                                   int32_t *ref_wp, double *ratio, int64_t n) {
  Map map = m->map;
  for (int64_t i = 0; i < n; i++) {
    map.init_reference_waypoint(x[i], y[i]);
    ref_wp[i] = map.reference_waypoint_id;
    for (int l = 0; l < 3; l++) ratio[i * 3 + l] = map.reference_waypoint_ratio[l];
  }
}

void ppref_lane_matching(ppref_map *m, const double *rx, const double *ry, const double *x,
                         const double *y, const double *vx, const double *vy, int32_t *ok,
                         int32_t *lane, int32_t *next_wp, double *s, double *d, double *vs,
                         double *vd, int64_t n) {
  Map map = m->map;
  for (int64_t i = 0; i < n; i++) {
    map.init_reference_waypoint(rx[i], ry[i]);
    int ln = 0, nwp = 0;
    double ss = 0, dd = 0, a = 0, b = 0;
    bool good = map.lane_matching(x[i], y[i], ss, dd, ln, &nwp);
    if (good) map.project_speed(Point(vx[i], vy[i]), nwp, &a, &b);
    ok[i] = good ? 1 : 0;
    lane[i] = good ? ln : -1;
    next_wp[i] = good ? nwp : 0;
    s[i] = good ? ss : 0;
    d[i] = good ? dd : 0;
    vs[i] = good ? a : 0;
    vd[i] = good ? b : 0;
  }
}

void ppref_get_lane_pos(ppref_map *m, const double *rx, const double *ry, const double *s,
                        const int32_t *lane, double *ox, double *oy, int32_t *owp, double *odist,
                        int64_t n) {
  Map map = m->map;
  for (int64_t i = 0; i < n; i++) {
    map.init_reference_waypoint(rx[i], ry[i]);
    int wp = 0;
    double dist = 0;
    Point p = map.get_lane_pos(s[i], lane[i], wp, dist);
    ox[i] = p.x;
    oy[i] = p.y;
    owp[i] = wp;
    odist[i] = dist;
  }
}

void ppref_spline(const double *kx, const double *ky, int32_t nk, const double *q, int32_t nq,
                  double *out, int64_t ns) {
  for (int64_t i = 0; i < ns; i++) {
    vector<double> xs(kx + i * nk, kx + (i + 1) * nk), ys(ky + i * nk, ky + (i + 1) * nk);
    tk::spline sp;
    sp.set_points(xs, ys);
    for (int j = 0; j < nq; j++) out[i * nq + j] = sp(q[i * nq + j]);
  }
}

void ppref_closest_waypoint(const double *x, const double *y, const double *mx, const double *my,
                            int32_t nwp, int32_t *out, int64_t n) {
  vector<double> vx(mx, mx + nwp), vy(my, my + nwp);
  for (int64_t i = 0; i < n; i++) out[i] = ClosestWaypoint(x[i], y[i], vx, vy);
}
void ppref_next_waypoint(const double *x, const double *y, const double *th, const double *mx,
                         const double *my, int32_t nwp, int32_t *out, int64_t n) {
  vector<double> vx(mx, mx + nwp), vy(my, my + nwp);
  for (int64_t i = 0; i < n; i++) out[i] = NextWaypoint(x[i], y[i], th[i], vx, vy);
}
void ppref_get_frenet(const double *x, const double *y, const double *th, const double *mx,
                      const double *my, int32_t nwp, double *os, double *od, int64_t n) {
  vector<double> vx(mx, mx + nwp), vy(my, my + nwp);
  for (int64_t i = 0; i < n; i++) {
    vector<double> r = getFrenet(x[i], y[i], th[i], vx, vy);
    os[i] = r[0];
    od[i] = r[1];
  }
}
void ppref_get_xy(const double *s, const double *d, const double *ms, const double *mx,
                  const double *my, int32_t nwp, double *ox, double *oy, int64_t n) {
  vector<double> vs(ms, ms + nwp), vx(mx, mx + nwp), vy(my, my + nwp);
  for (int64_t i = 0; i < n; i++) {
    vector<double> r = getXY(s[i], d[i], vs, vx, vy);
    ox[i] = r[0];
    oy[i] = r[1];
  }
}

// LaneChangePlanner::calculate_target_lane on explicit car lists:
// per problem i, cars j in [0,nc): id, s, vs, lane (others unused).
void ppref_lane_change(const int32_t *car_id, const double *car_s, const double *car_vs,
                       const int32_t *car_lane, int32_t nc, const int32_t *ego_lane,
                       const int32_t *target_lane, const double *ego_s, const double *ego_vs,
                       const double *dt0, int32_t *out, int64_t n) {
  for (int64_t i = 0; i < n; i++) {
    std::map<int, Car> cars;
    for (int j = 0; j < nc; j++) {
      Car c;
      memset(&c, 0, sizeof c);
      c.id = car_id[i * nc + j];
      c.s = car_s[i * nc + j];
      c.vs = car_vs[i * nc + j];
      c.lane = car_lane[i * nc + j];
      if (c.lane < 0) continue;
      cars[c.id] = c;
    }
    LaneChangePlanner p;
    out[i] = p.calculate_target_lane(cars, ego_lane[i], target_lane[i], ego_s[i], ego_vs[i],
                                     dt0[i]);
  }
}

// LimitSpeed::calculate + SpeedController::add_limit_breakpoint for one
// followed car: out_speed/out_time = controller target after the limit.
void ppref_limit_speed(const double *car_vx, const double *car_vy, const double *next_s,
                       const double *ego_s, const double *ego_speed, const double *ego_acc,
                       const int32_t *in_lane, double *ls_speed, double *ls_time,
                       double *sc_speed, double *sc_time, uint32_t *flags, int64_t n) {
  fLog = devnull();
  for (int64_t i = 0; i < n; i++) {
    t_flags = 0;
    Car c;
    memset(&c, 0, sizeof c);
    c.vx = car_vx[i];
    c.vy = car_vy[i];
    LimitSpeed ls;
    ls.calculate(c, next_s[i], ego_s[i], ego_speed[i], ego_acc[i], in_lane[i] != 0);
    SpeedController sc(ego_speed[i]);
    sc.add_limit_breakpoint(ls.target_speed, ls.target_time);
    ls_speed[i] = ls.target_speed;
    ls_time[i] = ls.target_time;
    sc_speed[i] = sc.target_speed;
    sc_time[i] = sc.target_time;
    flags[i] = t_flags;
  }
  fLog = NULL;
}

// TrajectoryBuilder::build on explicit inputs.
void ppref_trajectory_build(ppref_map *m, const int32_t *prev_n, const double *prev_x,
                            const double *prev_y, const double *ego_x, const double *ego_y,
                            const double *yaw, const int32_t *target_lane, const double *ego_d,
                            const double *ego_vd, const double *sc_start, const double *sc_target,
                            const double *sc_time, double *out_x, double *out_y, int32_t *out_n,
                            uint32_t *out_flags, int64_t n) {
  Map map = m->map;
  fLog = devnull();
  const int K = PP_PREV_KEEP;
  for (int64_t i = 0; i < n; i++) {
    t_flags = 0;
    vector<Point> prev;
    if (prev_n[i] >= K)
      for (int k = 0; k < K; k++) prev.push_back(Point(prev_x[i * K + k], prev_y[i * K + k]));
    double rx = prev.empty() ? ego_x[i] : prev[K - 1].x;
    double ry = prev.empty() ? ego_y[i] : prev[K - 1].y;
    map.init_reference_waypoint(rx, ry);
    SpeedController sc(sc_start[i]);
    sc.target_speed = sc_target[i];
    sc.target_time = sc_time[i];
    TrajectoryBuilder tb;
    vector<Point> tr = tb.build(prev, rx, ry, yaw[i], 0, target_lane[i], ego_d[i], ego_vd[i], map, sc);
    out_n[i] = (int32_t)tr.size();
    for (size_t k = 0; k < tr.size() && k < PP_PATH_LEN; k++) {
      out_x[i * PP_PATH_LEN + k] = tr[k].x;
      out_y[i * PP_PATH_LEN + k] = tr[k].y;
    }
    out_flags[i] = t_flags;
  }
  fLog = NULL;
}

// ---------------------------------------------------------------------------
// lambda harness: drive the untouched onMessage lambda.
// ---------------------------------------------------------------------------
struct LambdaJob {
  const pp_frames *in;
  int64_t n;
  double *ox, *oy;
  int32_t *on;
};
static LambdaJob *g_job = NULL;
struct LambdaDone {};

static void put(std::string &s, double v) {
  char buf[40];
  snprintf(buf, sizeof buf, "%.17g", v);
  s += buf;
}

}  // extern "C"

void uWS::Hub::run() {
  LambdaJob *job = g_job;
  const pp_frames *in = job->in;
  for (int64_t f = 0; f < job->n; f++) {
    std::string s = "42[\"telemetry\",{\"x\":";
    put(s, in->ego_x[f]);
    s += ",\"y\":";
    put(s, in->ego_y[f]);
    s += ",\"s\":0,\"d\":0,\"yaw\":";
    put(s, in->ego_yaw_deg[f]);
    s += ",\"speed\":";
    put(s, in->ego_speed_mph[f]);
    for (int axis = 0; axis < 2; axis++) {
      s += axis ? ",\"previous_path_y\":[" : ",\"previous_path_x\":[";
      const double *src = axis ? in->prev_y : in->prev_x;
      int pn = in->prev_n[f];
      for (int i = 0; i < pn; i++) {
        if (i) s += ",";
        // only the first 10 matter (src/main.cpp:1261-1268); pad with the last stored one
        int k = i < PP_PREV_KEEP ? i : PP_PREV_KEEP - 1;
        put(s, src[f * PP_PREV_KEEP + k]);
      }
      s += "]";
    }
    s += ",\"end_path_s\":0,\"end_path_d\":0,\"sensor_fusion\":[";
    int mc = in->max_cars;
    for (int j = 0; j < in->n_cars[f]; j++) {
      if (j) s += ",";
      s += "[";
      s += std::to_string(in->car_id[f * mc + j]);
      const double *arr[4] = {in->car_x, in->car_y, in->car_vx, in->car_vy};
      for (int k = 0; k < 4; k++) {
        s += ",";
        put(s, arr[k][f * mc + j]);
      }
      s += ",0,0]";
    }
    s += "]}]";
    std::vector<char> buf(s.begin(), s.end());
    buf.push_back(0);
    uWS::send_sink().last.clear();
    message_fn(uWS::WebSocket<uWS::SERVER>(), buf.data(), s.size(), uWS::OpCode::TEXT);
    // reply: 42["control",{"next_x":[...],"next_y":[...]}]
    const std::string &r = uWS::send_sink().last;
    size_t b = r.find('{');
    auto j = json::parse(r.substr(b, r.rfind('}') - b + 1));
    std::vector<double> nx = j["next_x"], ny = j["next_y"];
    job->on[f] = (int32_t)nx.size();
    for (size_t i = 0; i < nx.size() && i < PP_PATH_LEN; i++) {
      job->ox[f * PP_PATH_LEN + i] = nx[i];
      job->oy[f * PP_PATH_LEN + i] = ny[i];
    }
  }
  throw LambdaDone();  // ref_main has no return after h.run(): leave by unwinding
}

extern "C" {

// Feed `n` frames IN ORDER through one fresh instance of the reference's
// main(): its own persistent state (target_lane starting at 1, the
// std::map<int,Car>) carries from frame to frame, so in->target_lane_in is
// ignored.  `data_dir_parent` must be a directory whose ../data/highway_map.csv
// is the map (src/main.cpp:1167).
int ppref_lambda_sequence(const char *cwd, const pp_frames *in, int64_t n, double *out_x,
                          double *out_y, int32_t *out_n) {
  char old[4096];
  if (!getcwd(old, sizeof old)) return PP_E_IO;
  if (chdir(cwd) != 0) return PP_E_IO;
  LambdaJob job = {in, n, out_x, out_y, out_n};
  g_job = &job;
  int rc = PP_E_ARG;
  std::streambuf *keep = std::cout.rdbuf();
  std::ostringstream quiet;
  std::cout.rdbuf(quiet.rdbuf());  // "Listening to port" chatter
  try {
    ref_main();
  } catch (LambdaDone &) {
    rc = PP_OK;
  }
  std::cout.rdbuf(keep);
  g_job = NULL;
  if (chdir(old) != 0) return PP_E_IO;
  return rc;
}

}  // extern "C"
