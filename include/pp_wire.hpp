// pp_wire.hpp — the simulator wire format of the reference, socket-free.
//
//   has_data            hasData()                                   src/helpers.h:15-25
//   parse_telemetry     42["telemetry",{...}] -> pp::Frame          src/main.cpp:1217-1252,1297,1328-1334
//   control_message     pp::Plan -> 42["control",{"next_x":[..],"next_y":[..]}]   :1461-1466
//   manual_message      42["manual",{}]                             :1469-1471
//   TrajectoryLog       trajectory.log in the layout DrawLines.ipynb reads       :1198-1208,776-778,1044-1046
//
// Transport (uWebSockets) stays out of scope: these functions turn recorded or live message
// strings into frames for pp::Planner and plans back into reply strings.  The JSON reader is a
// minimal one for exactly this schema (objects, arrays, numbers, strings, null/true/false);
// numbers are printed with 15 significant digits like the reference's JSON library
// (src/json.hpp:6689).  Header-only, C++11.
#ifndef PP_B200_WIRE_HPP
#define PP_B200_WIRE_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "pp.hpp"

namespace pp {
namespace wire {

// src/helpers.h:15-25
inline std::string has_data(const std::string &s) {
  const size_t found_null = s.find("null");
  const size_t b1 = s.find_first_of("[");
  const size_t b2 = s.find_first_of("}");
  if (found_null != std::string::npos) return "";
  if (b1 != std::string::npos && b2 != std::string::npos) return s.substr(b1, b2 - b1 + 2);
  return "";
}

namespace detail {
struct Value {  // a JSON value of the telemetry schema
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  double num = 0;
  bool flag = false;
  std::string str;
  std::vector<Value> items;                               // Array
  std::vector<std::pair<std::string, Value>> members;     // Object
  const Value *find(const char *key) const {
    for (const auto &m : members)
      if (m.first == key) return &m.second;
    return nullptr;
  }
};
struct Reader {
  const char *p, *end;
  bool ok;
  int depth;  // nesting of the value being read; the schema needs 4 (untrusted socket input)
  Reader(const char *b, const char *e) : p(b), end(e), ok(true), depth(0) {}
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) {
      p++;
      return true;
    }
    return false;
  }
  std::string string() {
    std::string out;
    if (!eat('"')) {
      ok = false;
      return out;
    }
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) p++;  // the schema has no escapes that matter
      out += *p++;
    }
    if (p < end) p++;
    else ok = false;
    return out;
  }
  Value value() {
    Value v;
    ws();
    if (p >= end || depth > 16) {
      ok = false;
      return v;
    }
    if (*p == '{') {
      p++;
      depth++;
      v.kind = Value::Object;
      if (eat('}')) {
        depth--;
        return v;
      }
      do {
        std::string k = string();
        if (!eat(':')) ok = false;
        v.members.emplace_back(k, value());
      } while (ok && eat(','));
      if (!eat('}')) ok = false;
      depth--;
    } else if (*p == '[') {
      p++;
      depth++;
      v.kind = Value::Array;
      if (eat(']')) {
        depth--;
        return v;
      }
      do v.items.push_back(value());
      while (ok && eat(','));
      if (!eat(']')) ok = false;
      depth--;
    } else if (*p == '"') {
      v.kind = Value::String;
      v.str = string();
    } else if (end - p >= 4 && std::string(p, 4) == "null") {
      p += 4;
    } else if (end - p >= 4 && std::string(p, 4) == "true") {
      p += 4;
      v.kind = Value::Bool;
      v.flag = true;
    } else if (end - p >= 5 && std::string(p, 5) == "false") {
      p += 5;
      v.kind = Value::Bool;
    } else {
      char *q = nullptr;
      v.num = std::strtod(p, &q);
      if (q == p) ok = false;
      v.kind = Value::Number;
      p = q;
    }
    return v;
  }
};
inline bool numbers(const Value *v, std::vector<double> &out) {
  if (!v || v->kind != Value::Array) return false;
  out.clear();
  for (const Value &x : v->items) {
    if (x.kind != Value::Number) return false;
    out.push_back(x.num);
  }
  return true;
}
inline void put_number(std::string &s, double v) {
  if (!std::isfinite(v)) {  // nlohmann::json 3.0.0 prints a non-finite number as null
    s += "null";            // (src/json.hpp dump_float); "nan" / "inf" are not JSON
    return;
  }
  char buf[40];
  std::snprintf(buf, sizeof buf, "%.15g", v);
  s += buf;
  bool plain = true;
  for (const char *c = buf; *c; c++)
    if (*c == '.' || *c == 'e' || *c == 'n' || *c == 'i') plain = false;
  if (plain) s += ".0";
}
}  // namespace detail

enum MessageKind { NotAnEvent, Manual, Telemetry, Malformed };

// One websocket text message -> frame.  `target_lane` is the caller's persistent planner state
// (src/main.cpp:1195); it is copied into the frame.
inline MessageKind parse_telemetry(const std::string &msg, int target_lane, Frame &f) {
  if (msg.size() <= 2 || msg[0] != '4' || msg[1] != '2') return NotAnEvent;  // :1220
  const std::string s = has_data(msg);
  if (s.empty()) return Manual;  // :1468-1471
  detail::Reader r(s.data(), s.data() + s.size());
  const detail::Value j = r.value();
  if (!r.ok || j.kind != detail::Value::Array || j.items.size() < 2 ||
      j.items[0].kind != detail::Value::String)
    return Malformed;
  if (j.items[0].str != "telemetry") return NotAnEvent;
  const detail::Value &d = j.items[1];
  const detail::Value *x = d.find("x"), *y = d.find("y"), *yaw = d.find("yaw"), *sp = d.find("speed");
  if (!x || !y || !yaw || !sp) return Malformed;
  const detail::Value *scalars[4] = {x, y, yaw, sp};
  for (const detail::Value *v : scalars)
    if (v->kind != detail::Value::Number) return Malformed;  // (a non-number would read as 0)
  f = Frame();
  f.car_x = x->num;
  f.car_y = y->num;
  f.car_yaw = yaw->num;
  f.car_speed = sp->num;
  f.target_lane = target_lane;
  if (!detail::numbers(d.find("previous_path_x"), f.previous_path_x) ||
      !detail::numbers(d.find("previous_path_y"), f.previous_path_y) ||
      f.previous_path_x.size() != f.previous_path_y.size())
    return Malformed;
  const detail::Value *sf = d.find("sensor_fusion");
  if (!sf || sf->kind != detail::Value::Array) return Malformed;
  for (const detail::Value &row : sf->items) {  // [id, x, y, vx, vy, s, d] (:1328-1334)
    if (row.kind != detail::Value::Array || row.items.size() < 5) return Malformed;
    for (int k = 0; k < 5; k++)
      if (row.items[k].kind != detail::Value::Number) return Malformed;
    Car c;
    c.id = (int)row.items[0].num;
    c.x = row.items[1].num;
    c.y = row.items[2].num;
    c.vx = row.items[3].num;
    c.vy = row.items[4].num;
    f.sensor_fusion.push_back(c);
  }
  return Telemetry;
}

// The reference's cross-frame state (src/main.cpp:1194-1195): the persistent target_lane and
// the std::map<int,Car> sensor_fusion_cars, which OUTLIVES the frame.  Each message overwrites
// (or creates) the entries of the cars it lists — a repeated id keeps its last row — and a car
// that fails lane matching is erased (:1325-1340); an entry the message does not mention stays
// where it is, with the x, y, vx, vy AND the s, d, vs, vd, lane of its last sighting, and keeps
// taking part in LaneChangePlanner and the followed-car selection.  One Session per simulator
// connection reproduces that: frame_from() turns a message into the frame to plan (held-over
// cars marked frozen), update() folds the plan back into the state.
class Session {
 public:
  int target_lane = 1;        // :1195
  std::map<int, Car> cars;    // :1194

  MessageKind frame_from(const std::string &msg, Frame &f) {
    const MessageKind k = parse_telemetry(msg, target_lane, f);
    if (k != Telemetry) return k;
    std::set<int> listed;
    for (const Car &row : f.sensor_fusion) {  // :1325-1334, in message order
      Car &c = cars[row.id];
      c.id = row.id;
      c.x = row.x;
      c.y = row.y;
      c.vx = row.vx;
      c.vy = row.vy;
      listed.insert(row.id);
    }
    f.sensor_fusion.clear();
    for (const auto &kv : cars) {  // ascending id, like the reference's iteration
      Car c = kv.second;
      c.frozen = listed.count(kv.first) == 0;
      f.sensor_fusion.push_back(c);
    }
    return Telemetry;
  }

  void update(const Frame &f, const Plan &p) {
    target_lane = p.target_lane;  // :1355-1369
    for (size_t j = 0; j < f.sensor_fusion.size() && j < p.cars.size(); j++) {
      if (f.sensor_fusion[j].frozen) continue;
      const Car &r = p.cars[j];
      if (r.lane < 0) {  // :1336-1340
        cars.erase(r.id);
        continue;
      }
      Car &c = cars[r.id];
      c.s = r.s;
      c.d = r.d;
      c.vs = r.vs;
      c.vd = r.vd;
      c.lane = r.lane;
    }
  }
};

inline std::string control_message(const Plan &p) {  // :1461-1464
  std::string s = "42[\"control\",{\"next_x\":[";
  for (size_t i = 0; i < p.next_x.size(); i++) {
    if (i) s += ",";
    detail::put_number(s, p.next_x[i]);
  }
  s += "],\"next_y\":[";
  for (size_t i = 0; i < p.next_y.size(); i++) {
    if (i) s += ",";
    detail::put_number(s, p.next_y[i]);
  }
  s += "]}]";
  return s;
}
inline std::string manual_message() { return "42[\"manual\",{}]"; }  // :1470

// trajectory.log as the reference writes it when need_log is set: the map header once, then per
// frame, in the reference's order, "first control dist", the kept previous points, the builder's
// control points (:772-781) and the resulting trajectory (:1044-1046) — the arrays
// DrawLines.ipynb plots.  The free-text diagnostics of the individual branches are carried as
// bits on the flags line.
class TrajectoryLog {
 public:
  TrajectoryLog(const std::string &path, Map &map, const std::vector<double> &wx,
                const std::vector<double> &wy)
      : f_(std::fopen(path.c_str(), "wt")), map_(&map) {
    if (!f_) throw Error(PP_E_IO, "cannot open " + path);
    std::fprintf(f_, "wpmap=[");  // :1200-1202
    for (size_t i = 0; i < wx.size(); i++) std::fprintf(f_, "%s[%.4f,%.4f]", i == 0 ? "" : ",", wx[i], wy[i]);
    std::fprintf(f_, "]\n");
    for (int lane = 0; lane < PP_NUM_LANES; lane++) {  // :1203-1208
      std::fprintf(f_, "lane%d=[", lane);
      for (size_t i = 0; i < map.waypoints.size(); i++)
        std::fprintf(f_, "%s[%.4f,%.4f]", i == 0 ? "" : ",", map.waypoints[i].lane_center[lane].x,
                     map.waypoints[i].lane_center[lane].y);
      std::fprintf(f_, "]\n");
    }
  }
  ~TrajectoryLog() {
    if (f_) std::fclose(f_);
  }
  TrajectoryLog(const TrajectoryLog &) = delete;
  TrajectoryLog &operator=(const TrajectoryLog &) = delete;
  void frame(const Frame &in, const Plan &out) {
    std::fprintf(f_, "ego lane %d target lane %d\n", out.ego_lane, in.target_lane);  // :376
    std::fprintf(f_, "flags=0x%x target_lane=%d\n", out.flags, out.target_lane);
    const size_t keep = in.previous_path_x.size() >= PP_PREV_KEEP ? PP_PREV_KEEP : 0;  // :1261-1268
    const double pos_x = keep ? in.previous_path_x[keep - 1] : in.car_x;  // :583-600
    const double pos_y = keep ? in.previous_path_y[keep - 1] : in.car_y;
    const std::vector<Point> cp = TrajectoryBuilder::control_points(
        *map_, pos_x, pos_y, out.target_lane, out.ego_d, out.ego_vd, out.ego_speed);
    if (cp.size() > 1)  // :772-775 (the bare "]" closes a commented-out array there)
      std::fprintf(f_, "first control dist %.2f\n]\n", (Point(pos_x, pos_y) - cp[1]).length());
    std::fprintf(f_, "prev_trajectory=[");  // :776-778
    for (size_t i = 0; i < keep; i++)
      std::fprintf(f_, "%s[%.4f,%.4f]", i == 0 ? "" : ",", in.previous_path_x[i], in.previous_path_y[i]);
    std::fprintf(f_, "]\n");
    std::fprintf(f_, "control_points=[");  // :779-781
    for (size_t i = 0; i < cp.size(); i++)
      std::fprintf(f_, "%s[%.4f,%.4f]", i == 0 ? "" : ",", cp[i].x, cp[i].y);
    std::fprintf(f_, "]\n");
    std::fprintf(f_, "result=[");  // :1044-1046
    for (size_t i = 0; i < out.next_x.size(); i++)
      std::fprintf(f_, "%s[%.4f,%.4f]", i == 0 ? "" : ",", out.next_x[i], out.next_y[i]);
    std::fprintf(f_, "]\n");
    std::fflush(f_);  // :1458
  }

 private:
  std::FILE *f_;
  Map *map_;
};

}  // namespace wire
}  // namespace pp

#endif  // PP_B200_WIRE_HPP
