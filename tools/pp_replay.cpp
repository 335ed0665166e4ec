// pp_replay — replays recorded simulator sessions through the GPU planner, socket-free.
//
//   pp_replay --map data/highway_map.csv [--log trajectory.log] session1.txt [session2.txt ...] 
//
// Each session file holds one websocket text message per line, as the Udacity simulator sends
// them (42["telemetry",{...}]).  The sessions advance in lockstep: step k plans the k-th message
// of every session in ONE pp::Planner::plan call (a batch of #sessions frames), carries each
// session's cross-frame state (target_lane and the persistent car map, src/main.cpp:1194-1195:
// pp::wire::Session) and writes the reply the reference would send
// (42["control",{...}] / 42["manual",{}]) to <session>.out, one per line.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>

#include "pp_wire.hpp"

int main(int argc, char **argv) {
  std::string map_csv = "data/highway_map.csv", log_path;
  std::vector<std::string> files;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "--map" && i + 1 < argc) map_csv = argv[++i];
    else if (a == "--log" && i + 1 < argc) log_path = argv[++i];
    else files.push_back(a);
  }
  if (files.empty()) {
    std::fprintf(stderr, "usage: pp_replay --map highway_map.csv [--log trajectory.log] session.txt...\n");
    return 64;
  }
  try {
    pp::Map map;
    map.InitFromCsv(map_csv);
    pp::Planner planner(map);
    std::unique_ptr<pp::wire::TrajectoryLog> log;
    if (!log_path.empty()) {
      std::vector<double> wx, wy;
      for (const auto &w : map.waypoints) {
        wx.push_back(w.ref.x);
        wy.push_back(w.ref.y);
      }
      log.reset(new pp::wire::TrajectoryLog(log_path, map, wx, wy));
    }
    const size_t ns = files.size();
    std::vector<std::ifstream> in(ns);
    std::vector<std::ofstream> out(ns);
    for (size_t s = 0; s < ns; s++) {
      in[s].open(files[s]);
      out[s].open(files[s] + ".out");
      if (!in[s] || !out[s]) {
        std::fprintf(stderr, "cannot open %s(.out)\n", files[s].c_str());
        return 66;
      }
    }
    std::vector<pp::wire::Session> session(ns);  // :1194-1195
    long frames = 0, steps = 0;
    for (;;) {
      std::vector<pp::Frame> batch;
      std::vector<size_t> owner;
      bool any = false;
      for (size_t s = 0; s < ns; s++) {
        std::string line;
        if (!std::getline(in[s], line)) continue;
        any = true;
        pp::Frame f;
        switch (session[s].frame_from(line, f)) {
          case pp::wire::Telemetry:
            batch.push_back(f);
            owner.push_back(s);
            break;
          case pp::wire::Manual:
            out[s] << pp::wire::manual_message() << "\n";
            break;
          case pp::wire::Malformed:
            std::fprintf(stderr, "%s: malformed message skipped\n", files[s].c_str());
            out[s] << "\n";
            break;
          default:
            out[s] << "\n";  // not a websocket event: the reference sends nothing
        }
      }
      if (!any) break;
      if (!batch.empty()) {
        const std::vector<pp::Plan> plans = planner.plan(batch);
        for (size_t k = 0; k < plans.size(); k++) {
          const size_t s = owner[k];
          session[s].update(batch[k], plans[k]);
          out[s] << pp::wire::control_message(plans[k]) << "\n";
          if (log) log->frame(batch[k], plans[k]);
        }
        frames += (long)plans.size();
      }
      steps++;
    }
    std::fprintf(stderr, "pp_replay: %ld frames in %ld steps over %zu session(s)\n", frames, steps, ns);
  } catch (const pp::Error &e) {
    std::fprintf(stderr, "pp::Error %d: %s\n", e.code, e.what());
    return 2;
  }
  return 0;
}
