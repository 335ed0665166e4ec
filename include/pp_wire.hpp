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

#include <cstdio>
#include <cstdlib>
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
  Reader(const char *b, const char *e) : p(b), end(e), ok(true) {}
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
    if (p >= end) {
      ok = false;
      return v;
    }
    if (*p == '{') {
      p++;
      v.kind = Value::Object;
      if (eat('}')) return v;
      do {
        std::string k = string();
        if (!eat(':')) ok = false;
        v.members.emplace_back(k, value());
      } while (ok && eat(','));
      if (!eat('}')) ok = false;
    } else if (*p == '[') {
      p++;
      v.kind = Value::Array;
      if (eat(']')) return v;
      do v.items.push_back(value());
      while (ok && eat(','));
      if (!eat(']')) ok = false;
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
// frame the kept previous points and the resulting trajectory.  (control_points= and the
// free-text diagnostics of the reference are internal to its builder; the per-frame flags
// line carries the same information as bits.)
class TrajectoryLog {
 public:
  TrajectoryLog(const std::string &path, Map &map, const std::vector<double> &wx,
                const std::vector<double> &wy)
      : f_(std::fopen(path.c_str(), "wt")) {
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
    std::fprintf(f_, "prev_trajectory=[");  // :776-778
    for (size_t i = 0; i < keep; i++)
      std::fprintf(f_, "%s[%.4f,%.4f]", i == 0 ? "" : ",", in.previous_path_x[i], in.previous_path_y[i]);
    std::fprintf(f_, "]\n");
    std::fprintf(f_, "result=[");  // :1044-1046
    for (size_t i = 0; i < out.next_x.size(); i++)
      std::fprintf(f_, "%s[%.4f,%.4f]", i == 0 ? "" : ",", out.next_x[i], out.next_y[i]);
    std::fprintf(f_, "]\n");
    std::fflush(f_);  // :1458
  }

 private:
  std::FILE *f_;
};

}  // namespace wire
}  // namespace pp

#endif  // PP_B200_WIRE_HPP
