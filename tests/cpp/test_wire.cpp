// Host-side logic of include/pp_wire.hpp, no GPU needed: the simulator's wire format
// (src/main.cpp:1217-1252,1325-1340,1461-1471; src/helpers.h:15-25) and the cross-frame state a
// session carries (:1194-1195).  Exits 0 and prints "wire ok", or names the first failed check.
#include <cmath>
#include <cstdio>
#include <limits>
#include <string>

#include "pp_wire.hpp"

static int g_failed = 0;
#define CHECK(cond)                                                  \
  do {                                                               \
    if (!(cond)) {                                                   \
      std::printf("FAILED line %d: %s\n", __LINE__, #cond);          \
      g_failed++;                                                    \
    }                                                                \
  } while (0)

static std::string telemetry(const std::string &cars, const std::string &extra = "") {
  return "42[\"telemetry\",{\"x\":909.48,\"y\":1128.67,\"yaw\":0.5,\"speed\":21.25,\"s\":124.8,"
         "\"d\":6.16,\"previous_path_x\":[910.0,910.5,911.25],\"previous_path_y\":[1128.7,1128.8,"
         "1128.9],\"end_path_s\":0,\"end_path_d\":0,\"sensor_fusion\":[" + cars + "]" + extra + "}]";
}

int main() {
  using namespace pp::wire;
  pp::Frame f;
  // ---- framing (:1220, helpers.h:15-25)
  CHECK(parse_telemetry("2", 1, f) == NotAnEvent);
  CHECK(parse_telemetry("40", 1, f) == NotAnEvent);
  CHECK(parse_telemetry("42[\"telemetry\",null]", 1, f) == Manual);  // "null" anywhere: manual driving
  CHECK(parse_telemetry("42", 1, f) == NotAnEvent);
  CHECK(parse_telemetry("42[\"other\",{\"x\":1}]", 1, f) == NotAnEvent);
  CHECK(has_data("42[\"telemetry\",{\"x\":1}]") == "[\"telemetry\",{\"x\":1}]");
  CHECK(has_data("42[\"telemetry\",null]").empty());
  // ---- a well-formed message
  const std::string two = "[7,900.5,1120.25,20.0,0.5,130.0,2.0],[3,950.0,1130.0,18.5,-0.25,170.0,6.0]";
  CHECK(parse_telemetry(telemetry(two), 2, f) == Telemetry);
  CHECK(f.car_x == 909.48 && f.car_y == 1128.67 && f.car_yaw == 0.5 && f.car_speed == 21.25);
  CHECK(f.target_lane == 2);
  CHECK(f.previous_path_x.size() == 3 && f.previous_path_y.size() == 3);
  CHECK(f.previous_path_x[2] == 911.25 && f.previous_path_y[0] == 1128.7);
  CHECK(f.sensor_fusion.size() == 2);
  CHECK(f.sensor_fusion[0].id == 7 && f.sensor_fusion[0].x == 900.5 && f.sensor_fusion[0].vy == 0.5);
  CHECK(f.sensor_fusion[1].id == 3 && f.sensor_fusion[1].vx == 18.5 && f.sensor_fusion[1].vy == -0.25);
  // exponents, negative numbers, values the planner never reads (nested arrays, strings with
  // escapes, booleans)
  CHECK(parse_telemetry(telemetry("[1,-1.5e2,2E1,0,-0.0]", ",\"junk\":[1,[2,[3]]],\"b\":\"x\\\"y\",\"c\":true"), 0, f) ==
        Telemetry);
  CHECK(f.sensor_fusion.size() == 1 && f.sensor_fusion[0].x == -150.0 && f.sensor_fusion[0].y == 20.0);
  CHECK(parse_telemetry(telemetry(""), 1, f) == Telemetry && f.sensor_fusion.empty());
  // the reference cuts the payload at the FIRST '}' (hasData, src/helpers.h:20-23): an object
  // nested inside the telemetry object truncates it, there as here
  CHECK(parse_telemetry(telemetry("", ",\"o\":{\"a\":1}"), 1, f) == Malformed);
  // ---- malformed input is refused, not read as zeros
  CHECK(parse_telemetry(telemetry("[1,2,3,4]"), 1, f) == Malformed);                 // short car row
  CHECK(parse_telemetry(telemetry("[1,2,\"3\",4,5]"), 1, f) == Malformed);           // non-number
  CHECK(parse_telemetry("42[\"telemetry\",{\"x\":\"1\",\"y\":2,\"yaw\":3,\"speed\":4,"
                        "\"previous_path_x\":[],\"previous_path_y\":[],\"sensor_fusion\":[]}]", 1, f) == Malformed);
  CHECK(parse_telemetry("42[\"telemetry\",{\"x\":1,\"y\":2,\"yaw\":3,\"speed\":4,"
                        "\"previous_path_x\":[1,2],\"previous_path_y\":[1],\"sensor_fusion\":[]}]", 1, f) == Malformed);
  CHECK(parse_telemetry("42[\"telemetry\",{\"x\":1,\"y\":2,\"yaw\":3,"
                        "\"previous_path_x\":[],\"previous_path_y\":[],\"sensor_fusion\":[]}]", 1, f) == Malformed);
  CHECK(parse_telemetry("42[\"telemetry\",{\"x\":1,\"y\":2,\"yaw\":3,\"speed\":4,\"previous_path_x\":[],"
                        "\"previous_path_y\":[],\"sensor_fusion\":[[1,2,3,4,5]}]", 1, f) == Malformed);
  {  // nesting beyond the reader's limit: refused, the stack does not grow with the input
    std::string deep = "42[\"telemetry\",{\"x\":1,\"y\":2,\"yaw\":3,\"speed\":4,\"previous_path_x\":[],"
                       "\"previous_path_y\":[],\"sensor_fusion\":[],\"z\":";
    for (int i = 0; i < 4000; i++) deep += "[";
    for (int i = 0; i < 4000; i++) deep += "]";
    deep += "}]";
    CHECK(parse_telemetry(deep, 1, f) == Malformed);
  }
  // ---- replies (:1461-1471); numbers as nlohmann::json 3.0.0 prints them
  pp::Plan p;
  p.next_x = {1.0, 2.5, 1e21, 0.1 + 0.2};
  p.next_y = {-3.0, std::numeric_limits<double>::quiet_NaN(), std::numeric_limits<double>::infinity(), 1e-7};
  CHECK(control_message(p) ==
        "42[\"control\",{\"next_x\":[1.0,2.5,1e+21,0.3],\"next_y\":[-3.0,null,null,1e-07]}]");
  CHECK(manual_message() == "42[\"manual\",{}]");
  p.next_x.clear();
  p.next_y.clear();
  CHECK(control_message(p) == "42[\"control\",{\"next_x\":[],\"next_y\":[]}]");
  // ---- the persistent car map of a session (:1194, 1325-1340)
  Session s;
  CHECK(s.target_lane == 1 && s.cars.empty());
  CHECK(s.frame_from(telemetry(two), f) == Telemetry);
  CHECK(f.target_lane == 1);
  CHECK(f.sensor_fusion.size() == 2 && f.sensor_fusion[0].id == 3 && f.sensor_fusion[1].id == 7);  // ascending id
  CHECK(!f.sensor_fusion[0].frozen && !f.sensor_fusion[1].frozen);
  pp::Plan plan;  // what the planner reported for those two cars
  plan.target_lane = 0;
  plan.cars = f.sensor_fusion;
  plan.cars[0].lane = 1; plan.cars[0].s = 45.0; plan.cars[0].d = 6.0; plan.cars[0].vs = 18.0; plan.cars[0].vd = 0.1;
  plan.cars[1].lane = 0; plan.cars[1].s = 5.5; plan.cars[1].d = 2.0; plan.cars[1].vs = 20.0; plan.cars[1].vd = 0.0;
  s.update(f, plan);
  CHECK(s.target_lane == 0 && s.cars.size() == 2 && s.cars[3].s == 45.0 && s.cars[7].lane == 0);
  // car 7 drops out of the next message, car 9 appears, car 3 is listed twice (the last row wins)
  CHECK(s.frame_from(telemetry("[3,951.0,1130.5,18.5,0,0,0],[9,800.0,1100.0,22.0,0,0,0],[3,952.0,1131.0,18.0,0,0,0]"), f) ==
        Telemetry);
  CHECK(f.target_lane == 0);
  CHECK(f.sensor_fusion.size() == 3);
  CHECK(f.sensor_fusion[0].id == 3 && f.sensor_fusion[0].x == 952.0 && !f.sensor_fusion[0].frozen);
  CHECK(f.sensor_fusion[1].id == 7 && f.sensor_fusion[1].frozen);  // held over with its last sighting
  CHECK(f.sensor_fusion[1].x == 900.5 && f.sensor_fusion[1].s == 5.5 && f.sensor_fusion[1].lane == 0 &&
        f.sensor_fusion[1].vs == 20.0);
  CHECK(f.sensor_fusion[2].id == 9 && !f.sensor_fusion[2].frozen);
  // car 9 fails lane matching: erased (:1336-1340); the frozen car's entry is not touched
  plan.cars = f.sensor_fusion;
  plan.target_lane = 2;
  plan.cars[0].lane = 1; plan.cars[0].s = 46.0;
  plan.cars[1].lane = 2; plan.cars[1].s = -999.0;  // whatever the planner echoes for a frozen car
  plan.cars[2].lane = -1;
  s.update(f, plan);
  CHECK(s.target_lane == 2 && s.cars.size() == 2 && s.cars.count(9) == 0);
  CHECK(s.cars[3].s == 46.0 && s.cars[7].s == 5.5 && s.cars[7].lane == 0);
  // a manual-driving message leaves the state alone
  CHECK(s.frame_from("42[\"telemetry\",null]", f) == Manual && s.cars.size() == 2 && s.target_lane == 2);
  if (g_failed) {
    std::printf("%d check(s) failed\n", g_failed);
    return 1;
  }
  std::printf("wire ok\n");
  return 0;
}
