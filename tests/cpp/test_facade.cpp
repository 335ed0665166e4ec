// Exercises include/pp.hpp the way host code written against the reference
// would: the onMessage glue (src/main.cpp:1254-1457) restated on the façade's
// classes for frame A / frame B of SURVEY.md Appendix B, checked against the
// known answers recorded there from the compiled reference, and against
// Planner::plan (the batched entry) on the same frames.
//
//   g++ -std=c++11 -I include tests/cpp/test_facade.cpp -L <pkg> -lpp_b200 -o test_facade
//   ./test_facade data/highway_map.csv
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "pp.hpp"

static int g_fail = 0;
#define EXPECT(cond)                                                     \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
      g_fail++;                                                          \
    }                                                                    \
  } while (0)
static bool close_rel(double a, double b, double rel = 1e-9, double abs_ = 1e-6) {
  return std::fabs(a - b) <= abs_ && std::fabs(a - b) <= rel * std::fmax(std::fabs(b), 1.0);
}

// The planning step of onMessage on the façade classes (one frame).
static std::vector<pp::Point> plan_with_classes(pp::Map &map, const pp::Frame &f, int &target_lane,
                                                double *ego_d_out = nullptr) {
  using namespace pp;
  double ego_x = f.car_x, ego_y = f.car_y, ego_yaw = f.car_yaw;
  double ego_speed = f.car_speed / 2.237, ego_acc = 0;  // :1239
  Point speed_vector;
  double delta_t0 = 0;
  std::vector<Point> prev_trajectory;
  const int keep = 10;
  if ((int)f.previous_path_x.size() >= keep) {  // :1261-1282
    for (int i = 0; i < keep; i++) prev_trajectory.push_back(Point(f.previous_path_x[i], f.previous_path_y[i]));
    const Point p1 = prev_trajectory[keep - 3], p2 = prev_trajectory[keep - 2], p3 = prev_trajectory[keep - 1];
    const double v2 = (p2 - p1).length(), v3 = (p3 - p2).length();
    ego_acc = (v3 - v2) * 50;
    ego_speed = v3 * 50;
    speed_vector = Point((p3.x - p2.x) * 50, (p3.y - p2.y) * 50);
    ego_x = p3.x;
    ego_y = p3.y;
    delta_t0 = keep / 50.0;
  }
  map.init_reference_waypoint(ego_x, ego_y);  // :1299
  int ego_lane = 0;
  double ego_s = 0, ego_d = 0, ego_vs = 0, ego_vd = 0;
  if (!map.lane_matching(ego_x, ego_y, ego_s, ego_d, ego_lane)) ego_s = ego_d = ego_lane = 0;
  map.project_speed(speed_vector, map.reference_waypoint_id, &ego_vs, &ego_vd);  // :1313
  if (ego_acc > 8) ego_acc = 8;
  if (ego_acc < -8) ego_acc = -8;
  std::map<int, Car> cars;  // :1325-1350
  for (const Car &c0 : f.sensor_fusion) {
    Car c = c0;
    if (map.match_car(c)) cars[c.id] = c;
  }
  LaneChangePlanner lcp;  // :1355
  int new_target = lcp.calculate_target_lane(cars, ego_lane, target_lane, ego_s, ego_vs, delta_t0);
  if (new_target != ego_lane) {  // :1358-1369
    const double d_target = map.get_lane_center_offset(new_target);
    if (std::fabs(ego_vd * 1.0 + ego_d - d_target) > 6.0) new_target = ego_lane;
  }
  target_lane = new_target;
  int next_car_id = -1, next_car_tl = -1;  // :1383-1411
  double next_s = 0, next_s_tl = 0;
  const double d_target = map.get_lane_center_offset(target_lane);
  for (auto &kv : cars) {
    const double s0 = kv.second.predicted_s(delta_t0), d0 = kv.second.predicted_d(delta_t0);
    if (s0 > ego_s && std::fabs(d0 - ego_d) < 3 && (next_car_id == -1 || next_s > s0)) {
      next_s = s0;
      next_car_id = kv.first;
    }
    if (s0 >= ego_s - 4.5 - 2 && std::fabs(d0 - d_target) < 3 && (next_car_tl == -1 || next_s_tl > s0)) {
      next_s_tl = s0;
      next_car_tl = kv.first;
    }
  }
  if (next_car_tl == next_car_id) next_car_tl = -1;
  SpeedController sc(ego_speed);  // :1422-1438
  if (next_car_id != -1) {
    LimitSpeed ls;
    ls.calculate(cars[next_car_id], next_s, ego_s, ego_speed, ego_acc, true);
    sc.add_limit_breakpoint(ls.target_speed, ls.target_time);
  }
  if (next_car_tl != -1) {
    LimitSpeed ls;
    ls.calculate(cars[next_car_tl], next_s_tl, ego_s, ego_speed, ego_acc, false);
    sc.add_limit_breakpoint(ls.target_speed, ls.target_time);
  }
  if (ego_d_out) *ego_d_out = ego_d;
  TrajectoryBuilder tb;  // :1446-1448
  return tb.build(prev_trajectory, ego_x, ego_y, ego_yaw, ego_lane, target_lane, ego_d, ego_vd, map, sc);
}

int main(int argc, char **argv) {
  const char *csv = argc > 1 ? argv[1] : "data/highway_map.csv";
  try {
    pp::Map map;
    map.InitFromCsv(csv);
    EXPECT(map.waypoints.size() == 181);
    // Map::Init
    EXPECT(map.waypoints[0].nx == -0.026938559669005456 && map.waypoints[0].ny == -0.99963709115006305);
    EXPECT(map.waypoints[0].lane_center[0].x == 784.55226438455873);
    EXPECT(map.waypoints[180].lane_center[0].y == 1134.426400317795);

    // distancesq_pt_seg KATs
    double rn, rd, sn;
    EXPECT(pp::distancesq_pt_seg(pp::Point(1, 2), pp::Point(0, 0), pp::Point(4, 0), rn, rd, sn) == 4 &&
           rn == 4 && rd == 16 && sn == -8);
    EXPECT(pp::distancesq_pt_seg(pp::Point(-0.2, 2), pp::Point(0, 0), pp::Point(4, 0), rn, rd, sn) == 4 &&
           rn == -0.8);
    EXPECT(pp::distancesq_pt_seg(pp::Point(-1, 2), pp::Point(0, 0), pp::Point(4, 0), rn, rd, sn) == 5 && rn == 0);

    // tk::spline KAT
    pp::tk::spline sp;
    sp.set_points({-3, -1, 0, 2, 5, 9}, {0.5, 0.1, 0, -0.2, 0.4, 1.5});
    EXPECT(sp(-4) == 0.73810308307837136 && sp(0.5) == -0.062533463756177937);
    EXPECT(sp(4) == 0.11692947360163186 && sp(10) == 1.7677983054836433);

    // SpeedController
    pp::SpeedController sc0(0);
    EXPECT(sc0.target_speed == 22.199999999999999 && sc0.target_time == 4.4399999999999995);
    EXPECT(close_rel(sc0.get_speed(2.22), 11.1));

    // frame A (cold start) and frame B (10 reused points)
    pp::Frame a;
    a.car_x = 909.48;
    a.car_y = 1128.67;
    a.car_yaw = 0;
    a.car_speed = 0;
    a.target_lane = 1;
    const double cars[3][5] = {{0, 1000, 1130, 15, 0.1}, {1, 950, 1126, 14, 0}, {2, 880, 1124.8, 20, 0}};
    for (auto &c : cars) {
      pp::Car k;
      k.id = (int)c[0];
      k.x = c[1];
      k.y = c[2];
      k.vx = c[3];
      k.vy = c[4];
      a.sensor_fusion.push_back(k);
    }
    int tl = 1;
    double ego_d = 0;
    std::vector<pp::Point> pa = plan_with_classes(map, a, tl, &ego_d);
    EXPECT(map.reference_waypoint_id == 5);
    EXPECT(map.reference_waypoint_ratio[0] == 0.13937061012750163);
    EXPECT(ego_d == 6.1660676684929081);
    EXPECT(tl == 1 && pa.size() == 50);
    EXPECT(close_rel(pa[0].x, 909.48199898938844) && close_rel(pa[0].y, 1128.670063572202));
    EXPECT(close_rel(pa[49].x, 912.02874432915644) && close_rel(pa[49].y, 1128.7500091916932));

    pp::Planner planner(map);
    pp::Plan qa = planner.plan(a);
    EXPECT(qa.next_x.size() == 50 && qa.target_lane == 1 && qa.ego_lane == 1 && qa.ref_wp == 5);
    for (int i = 0; i < 50; i++) EXPECT(qa.next_x[i] == pa[i].x && qa.next_y[i] == pa[i].y);

    pp::Frame b = a;  // previous path = A's points 3..49
    for (int i = 3; i < 50; i++) {
      b.previous_path_x.push_back(pa[i].x);
      b.previous_path_y.push_back(pa[i].y);
    }
    b.car_x = pa[2].x;
    b.car_y = pa[2].y;
    std::vector<pp::Point> pb = plan_with_classes(map, b, tl, &ego_d);
    EXPECT(pb.size() == 50 && tl == 1);
    EXPECT(close_rel(ego_d, 6.1618973809581528, 1e-12, 1e-9));
    EXPECT(close_rel(pb[10].x, 909.68989390670026) && close_rel(pb[49].x, 912.3406106400505));
    EXPECT(close_rel(pb[49].y, 1128.7591544107854));
    pp::Plan qb = planner.plan(b);
    for (int i = 0; i < 50; i++) EXPECT(qb.next_x[i] == pb[i].x && qb.next_y[i] == pb[i].y);
    EXPECT(close_rel(qb.ego_speed, 1.2999999999988825, 1e-9, 1e-9));

    // candidate sweep on frame B: the winner is a real candidate with a full trajectory, and
    // no candidate scores below it
    pp::Planner::Swept sw = planner.sweep(b);
    EXPECT(sw.best >= 0 && sw.best < PP_SWEEP_CANDS && sw.scores.size() == PP_SWEEP_CANDS);
    EXPECT(sw.next_x.size() == 50 && sw.score < PP_SWEEP_BAD && sw.score == sw.scores[sw.best]);
    for (double v : sw.scores) EXPECT(!(v < sw.score));
    EXPECT(sw.next_x[9] == b.previous_path_x[9] && sw.lane() >= 0 && sw.lane() <= 2);

    // closed-loop rollouts: 64 vehicles, 30 ticks — everybody moves, paths stay full
    pp::Rollouts ro(map, 64, 12, 7);
    const pp::Rollouts::Ego e0 = ro.ego();
    ro.run(30, 2);
    const pp::Rollouts::Ego e1 = ro.ego();
    EXPECT(e1.tick == 30);
    for (int i = 0; i < 64; i++) {
      EXPECT(pp::distance(e0.x[i], e0.y[i], e1.x[i], e1.y[i]) > 0.05);
      EXPECT(e1.path_n[i] == 48 && e1.target_lane[i] >= 0 && e1.target_lane[i] <= 2);
    }

    // starter helpers: a round trip inside the documented domain
    std::vector<double> mx, my, ms;
    double s_acc = 0;
    for (size_t i = 0; i < map.waypoints.size(); i++) {
      mx.push_back(map.waypoints[i].ref.x);
      my.push_back(map.waypoints[i].ref.y);
      if (i) s_acc += pp::distance(mx[i - 1], my[i - 1], mx[i], my[i]);
      ms.push_back(s_acc);
    }
    std::vector<double> xy = pp::getXY(100.0, 6.0, ms, mx, my);
    EXPECT(pp::ClosestWaypoint(xy[0], xy[1], mx, my) >= 2 && pp::ClosestWaypoint(xy[0], xy[1], mx, my) <= 4);
  } catch (const pp::Error &e) {
    std::printf("pp::Error %d: %s\n", e.code, e.what());
    return 2;
  }
  if (g_fail) {
    std::printf("%d expectation(s) failed\n", g_fail);
    return 1;
  }
  std::printf("facade ok\n");
  return 0;
}
