// pp_multi — the multi-GPU job of BASELINE configs[1] / configs[4] as ONE host program, one
// thread per device, through the C ABI alone (plain C++11, no CUDA headers).
//
//   pp_multi <highway_map.csv> [--gpus G] [--frames N] [--cars C] [--steps K] [--chunk F]
//            [--seed S] [--check]
//
// The N frames of one global counter-based stream are cut into contiguous shards, device g
// plans [g N / G, (g + 1) N / G).  Every device generates its shard in HBM
// (pp_synth_frames_dev), in chunks of at most F frames that all stay resident, so the timed
// region starts with the inputs in HBM.  A step = every device plans its shard
// (pp_plan_stats_batch per chunk); after the K steps the per-device int64 statistics and f64
// minima / maxima are reduced ONCE with pp_stats_reduce (NCCL; the only collective).  Timing
// follows SURVEY §8d "Multi-GPU timing": all devices synchronised, host wall clock from the
// first launch to the completion of the reduction on every device.
//
// --check: the same N frames are then planned on device 0 alone and the reduced statistics
// must be identical, bit for bit (exit code 1 otherwise).  Strong scaling = run with --gpus 1
// and --gpus G at the same N and divide.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "pp.h"

namespace {

struct Chunk {
  int64_t first = 0, n = 0;
  pp_frames in;
  pp_plans out;
};

#define CHECK(call)                                                                     \
  do {                                                                                  \
    const int rc_ = (call);                                                             \
    if (rc_ != PP_OK) {                                                                 \
      std::fprintf(stderr, "pp_multi: %s failed: %s %s\n", #call, pp_strerror(rc_),     \
                   rc_ == PP_E_CUDA ? pp_last_cuda_error() : "");                       \
      std::exit(2);                                                                     \
    }                                                                                   \
  } while (0)

template <class T>
T *dev_array(size_t count) {
  void *p = nullptr;
  CHECK(pp_dev_alloc(&p, count * sizeof(T)));
  return static_cast<T *>(p);
}

Chunk make_chunk(int64_t first, int64_t n, int cars) {
  Chunk c;
  c.first = first;
  c.n = n;
  const size_t N = (size_t)n, mc = (size_t)(cars > 0 ? cars : 1);
  std::memset(&c.in, 0, sizeof c.in);
  std::memset(&c.out, 0, sizeof c.out);
  c.in.ego_x = dev_array<double>(N);
  c.in.ego_y = dev_array<double>(N);
  c.in.ego_yaw_deg = dev_array<double>(N);
  c.in.ego_speed_mph = dev_array<double>(N);
  c.in.prev_n = dev_array<int32_t>(N);
  c.in.prev_x = dev_array<double>(N * PP_PREV_KEEP);
  c.in.prev_y = dev_array<double>(N * PP_PREV_KEEP);
  c.in.target_lane_in = dev_array<int32_t>(N);
  c.in.n_cars = dev_array<int32_t>(N);
  c.in.car_id = dev_array<int32_t>(N * mc);
  c.in.car_x = dev_array<double>(N * mc);
  c.in.car_y = dev_array<double>(N * mc);
  c.in.car_vx = dev_array<double>(N * mc);
  c.in.car_vy = dev_array<double>(N * mc);
  c.in.max_cars = (int32_t)mc;
  c.out.next_x = dev_array<double>(N * PP_PATH_LEN);
  c.out.next_y = dev_array<double>(N * PP_PATH_LEN);
  c.out.n_points = dev_array<int32_t>(N);
  c.out.ego_lane = dev_array<int32_t>(N);
  c.out.ref_wp = dev_array<int32_t>(N);
  c.out.target_lane = dev_array<int32_t>(N);
  c.out.flags = dev_array<uint32_t>(N);
  c.out.ego_speed = dev_array<double>(N);
  c.out.target_speed = dev_array<double>(N);
  return c;
}

struct Device {
  int id = 0;
  pp_map *map = nullptr;
  void *stream = nullptr;
  std::vector<Chunk> chunks;
  int64_t *stats_dev = nullptr, *chunk_stats_dev = nullptr;
  double *fstats_dev = nullptr, *chunk_fstats_dev = nullptr;
  std::vector<int64_t> stats;    // this device's totals of the last step
  std::vector<double> fstats;
};

// one pass over a device's shard: plans every chunk, leaves the totals in d.stats / d.fstats
void plan_shard(Device &d, const pp_config &cfg) {
  d.stats.assign(PP_STATS_LEN, 0);
  d.fstats.assign(PP_FSTATS_LEN, 0.0);
  for (int i = 0; i < PP_FSTATS_LEN; i++) d.fstats[i] = i < PP_FSTAT_NMIN ? HUGE_VAL : -HUGE_VAL;
  std::vector<int64_t> hs(PP_STATS_LEN);
  std::vector<double> hf(PP_FSTATS_LEN);
  for (Chunk &c : d.chunks) {
    CHECK(pp_plan_stats_batch(d.map, &cfg, &c.in, &c.out, c.n, d.chunk_stats_dev, d.stream));
    CHECK(pp_fstats_batch(&c.out, c.n, d.chunk_fstats_dev, d.stream));
    CHECK(pp_stream_sync(d.stream));
    CHECK(pp_dev_download(hs.data(), d.chunk_stats_dev, hs.size() * sizeof(int64_t)));
    CHECK(pp_dev_download(hf.data(), d.chunk_fstats_dev, hf.size() * sizeof(double)));
    for (int i = 0; i < PP_STATS_LEN; i++) d.stats[i] += hs[i];
    for (int i = 0; i < PP_FSTATS_LEN; i++)
      d.fstats[i] = i < PP_FSTAT_NMIN ? (hf[i] < d.fstats[i] ? hf[i] : d.fstats[i])
                                      : (hf[i] > d.fstats[i] ? hf[i] : d.fstats[i]);
  }
}

struct Barrier {  // (C++11: no std::barrier)
  std::atomic<int> count{0}, gen{0};
  int n;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    const int g = gen.load();
    if (count.fetch_add(1) + 1 == n) {
      count.store(0);
      gen.fetch_add(1);
    } else {
      while (gen.load() == g) std::this_thread::yield();
    }
  }
};

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: pp_multi <highway_map.csv> [--gpus G] [--frames N] [--cars C] "
                         "[--steps K] [--chunk F] [--seed S] [--check]\n");
    return 2;
  }
  const char *csv = argv[1];
  int gpus = 0, cars = 12, steps = 5;
  int64_t frames = 1 << 20, chunk = 1 << 21;
  uint64_t seed = 0x5EED;
  bool check = false;
  for (int i = 2; i < argc; i++) {
    const std::string a = argv[i];
    auto val = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
    if (a == "--gpus") gpus = std::atoi(val());
    else if (a == "--frames") frames = std::atoll(val());
    else if (a == "--cars") cars = std::atoi(val());
    else if (a == "--steps") steps = std::atoi(val());
    else if (a == "--chunk") chunk = std::atoll(val());
    else if (a == "--seed") seed = std::strtoull(val(), nullptr, 0);
    else if (a == "--check") check = true;
  }
  CHECK(pp_init());
  const int have = pp_device_count();
  if (have <= 0) {
    std::fprintf(stderr, "pp_multi: no CUDA device (%s): there is no CPU planning path\n",
                 pp_last_cuda_error());
    return 2;
  }
  const int G = gpus > 0 && gpus < have ? gpus : have;
  pp_config cfg;
  CHECK(pp_config_default(&cfg));

  std::vector<Device> dev(G);
  std::vector<void *> comms(G, nullptr);
  {
    std::vector<int> ids(G);
    for (int g = 0; g < G; g++) ids[g] = g;
    CHECK(pp_comm_init_all(G, ids.data(), comms.data()));
  }
  Barrier bar(G + 1);
  std::vector<double> t_step(G, 0.0);
  std::vector<std::thread> th;
  for (int g = 0; g < G; g++) {
    th.emplace_back([&, g]() {
      Device &d = dev[g];
      d.id = g;
      CHECK(pp_dev_set(g));
      CHECK(pp_map_create_from_csv(csv, &d.map));
      CHECK(pp_stream_create(&d.stream));
      d.stats_dev = dev_array<int64_t>(PP_STATS_LEN);
      d.chunk_stats_dev = dev_array<int64_t>(PP_STATS_LEN);
      d.fstats_dev = dev_array<double>(PP_FSTATS_LEN);
      d.chunk_fstats_dev = dev_array<double>(PP_FSTATS_LEN);
      const int64_t lo = g * frames / G, hi = (g + 1) * frames / G;
      for (int64_t f = lo; f < hi; f += chunk) {
        const int64_t n = hi - f < chunk ? hi - f : chunk;
        d.chunks.push_back(make_chunk(f, n, cars));
        Chunk &c = d.chunks.back();
        CHECK(pp_synth_frames_dev(d.map, seed, c.first, c.n, cars, 20, &c.in, d.stream));
      }
      CHECK(pp_stream_sync(d.stream));
      plan_shard(d, cfg);  // warm-up
      bar.wait();          // (the main thread warms the communicators up: NCCL sets its
      bar.wait();          //  channels up at the first collective, about a second)
      CHECK(pp_stream_sync(d.stream));
      bar.wait();          // ---- every device ready, inputs resident: the clock starts
      for (int k = 0; k < steps; k++) plan_shard(d, cfg);
      CHECK(pp_dev_upload(d.stats_dev, d.stats.data(), PP_STATS_LEN * sizeof(int64_t)));
      CHECK(pp_dev_upload(d.fstats_dev, d.fstats.data(), PP_FSTATS_LEN * sizeof(double)));
      bar.wait();  // the reduction is issued for all devices by the main thread (one NCCL group)
      bar.wait();
      CHECK(pp_stream_sync(d.stream));
      CHECK(pp_dev_download(d.stats.data(), d.stats_dev, PP_STATS_LEN * sizeof(int64_t)));
      CHECK(pp_dev_download(d.fstats.data(), d.fstats_dev, PP_FSTATS_LEN * sizeof(double)));
      bar.wait();  // ---- the clock stops
    });
  }
  auto reduce_all = [&]() {
    CHECK(pp_comm_group_begin());
    for (int g = 0; g < G; g++) {
      CHECK(pp_dev_set(g));
      CHECK(pp_stats_reduce(comms[g], dev[g].stats_dev, dev[g].fstats_dev, dev[g].stream));
    }
    CHECK(pp_comm_group_end());
  };
  bar.wait();
  reduce_all();  // warm-up (values are overwritten before the timed reduction)
  bar.wait();
  bar.wait();
  const auto t0 = std::chrono::steady_clock::now();
  bar.wait();
  const auto t_plan = std::chrono::steady_clock::now();
  reduce_all();
  bar.wait();
  bar.wait();
  const auto t1 = std::chrono::steady_clock::now();
  for (auto &t : th) t.join();
  const double secs = std::chrono::duration<double>(t1 - t0).count();
  const double reduce_ms = 1e3 * std::chrono::duration<double>(t1 - t_plan).count();
  for (int g = 1; g < G; g++)
    if (dev[g].stats != dev[0].stats || dev[g].fstats != dev[0].fstats) {
      std::fprintf(stderr, "pp_multi: device %d holds a different reduced vector\n", g);
      return 1;
    }
  int rc = 0;
  if (check && G > 1) {  // the same frames on device 0 alone
    CHECK(pp_dev_set(0));
    Device one;
    one.map = dev[0].map;
    one.stream = dev[0].stream;
    one.chunk_stats_dev = dev[0].chunk_stats_dev;
    one.chunk_fstats_dev = dev[0].chunk_fstats_dev;
    std::vector<int64_t> tot(PP_STATS_LEN, 0);
    std::vector<double> ftot(PP_FSTATS_LEN);
    for (int i = 0; i < PP_FSTATS_LEN; i++) ftot[i] = i < PP_FSTAT_NMIN ? HUGE_VAL : -HUGE_VAL;
    Chunk c = make_chunk(0, chunk < frames ? chunk : frames, cars);
    for (int64_t f = 0; f < frames; f += c.n) {
      const int64_t n = frames - f < c.n ? frames - f : c.n;
      Chunk view = c;
      view.first = f;
      view.n = n;
      CHECK(pp_synth_frames_dev(one.map, seed, f, n, cars, 20, &view.in, one.stream));
      one.chunks.assign(1, view);
      plan_shard(one, cfg);
      for (int i = 0; i < PP_STATS_LEN; i++) tot[i] += one.stats[i];
      for (int i = 0; i < PP_FSTATS_LEN; i++)
        ftot[i] = i < PP_FSTAT_NMIN ? (one.fstats[i] < ftot[i] ? one.fstats[i] : ftot[i])
                                    : (one.fstats[i] > ftot[i] ? one.fstats[i] : ftot[i]);
    }
    // (plan_shard starts its totals afresh, so the timed run's vector is that of ONE pass)
    const bool same = tot == dev[0].stats && ftot == dev[0].fstats;
    std::printf("check: %d-device reduced statistics %s the single-device statistics\n", G,
                same ? "==" : "!=");
    if (!same) rc = 1;
  }
  std::printf("{\"tool\": \"pp_multi\", \"gpus\": %d, \"frames\": %lld, \"cars\": %d, \"steps\": %d, "
              "\"seconds\": %.6f, \"frames_per_s\": %.1f, \"ms_per_step\": %.4f, "
              "\"reduce_ms\": %.4f, \"stat_frames\": %lld, \"stat_points\": %lld, "
              "\"max_acc\": %.17g}\n",
              G, (long long)frames, cars, steps, secs, (double)frames * steps / secs,
              1e3 * secs / steps, reduce_ms, (long long)dev[0].stats[PP_STAT_FRAMES],
              (long long)dev[0].stats[PP_STAT_POINTS], dev[0].fstats[PP_FSTAT_MAX_ACC]);
  return rc;
}
