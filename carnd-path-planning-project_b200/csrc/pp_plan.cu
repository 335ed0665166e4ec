// pp_plan.cu — the fused planning kernels and pp_plan_batch / pp_stats_batch.
//
// One kernel launch plans a whole batch: every stage of the reference's
// per-frame step (src/main.cpp:1254-1457) runs back to back in registers, the
// map table sits in shared memory, and nothing intermediate touches HBM.
//
// Mapping: the step is FP64-latency/issue bound, not HBM bound
// (1,520 algorithmic bytes against ~35k dependent FP64 instructions per
// frame), and ~85 % of those instructions sit in serial recurrences (segment
// walks, the 40-step emission loop).  The throughput kernel therefore gives
// every LANE its own frame, so the 32 lanes of a warp run 32 independent
// recurrences; the blocks are persistent (grid = multiple of the SM count)
// and stride over the batch.
#include <cuda_runtime.h>

#include <cstdio>

#include "pp_device.cuh"
#include "pp_internal.h"

namespace {

using namespace ppd;

constexpr int kBlock = 128;

// Stage A: ego state, src/main.cpp:1254-1282.
struct Ego {
  double x, y, speed, acc, svx, svy, dt0;
  int nprev;
};

PPD_INLINE Ego ego_state(const pp_frames &in, const pp_config &cfg, int64_t f, uint32_t &flags) {
  Ego e;
  e.x = in.ego_x[f];
  e.y = in.ego_y[f];
  e.speed = in.ego_speed_mph[f];
  e.speed /= 2.237;  // :1239
  e.acc = 0;
  e.svx = 0;
  e.svy = 0;
  e.dt0 = 0;
  e.nprev = 0;
  if (in.prev_n[f] >= PP_PREV_KEEP) {  // :1261
    const double *px = in.prev_x + f * PP_PREV_KEEP;
    const double *py = in.prev_y + f * PP_PREV_KEEP;
    const double p7x = px[7], p7y = py[7], p8x = px[8], p8y = py[8], p9x = px[9], p9y = py[9];
    const double v2 = vlen(p8x - p7x, p8y - p7y);
    e.svx = p9x - p8x;
    e.svy = p9y - p8y;
    const double v3 = vlen(e.svx, e.svy);
    e.acc = (v3 - v2) * 50;
    e.speed = v3 * 50;
    e.svx *= 50;
    e.svy *= 50;
    e.x = p9x;
    e.y = p9y;
    e.dt0 = PP_PREV_KEEP / 50.0;
    e.nprev = PP_PREV_KEEP;
  } else {
    flags |= PP_F_COLD_START;
  }
  // the clamp of :1319-1320 is applied by the caller after project_speed (it
  // only feeds LimitSpeed)
  (void)cfg;
  return e;
}

// Followed-car candidate: (s0, id) lexicographic minimum, == the reference's
// "first car in ascending id order with strictly smaller s0" (:1395,1404).
struct Cand {
  double s0;
  int id, j;
};
PPD_INLINE void cand_init(Cand &c) {
  c.s0 = 0;
  c.id = -1;
  c.j = -1;
}
PPD_INLINE void cand_offer(Cand &c, double s0, int id, int j) {
  if (c.id == -1 || c.s0 > s0 || (c.s0 == s0 && id < c.id)) {
    c.s0 = s0;
    c.id = id;
    c.j = j;
  }
}

// The whole planning step for frame f, executed by ONE thread.
__device__ void plan_frame(const MapView &m, const pp_config &cfg, const pp_frames &in,
                           const pp_plans &out, int64_t f) {
  uint32_t flags = 0;
  Ego e = ego_state(in, cfg, f, flags);

  RefState rs;
  init_reference(m, e.x, e.y, rs);  // :1299

  Match em = lane_match(m, rs, e.x, e.y);  // :1302-1307
  if (!em.ok) {
    flags |= PP_F_EGO_MATCH_FAIL;
    em.s = 0;
    em.d = 0;
    em.lane = 0;
  }
  double evs, evd;
  project_speed(m, e.svx, e.svy, rs.wp, evs, evd);  // :1313
  if (e.acc > cfg.maximum_acc) e.acc = cfg.maximum_acc;  // :1319-1320
  if (e.acc < -cfg.maximum_acc) e.acc = -cfg.maximum_acc;

  // ---- sensor fusion, one streaming pass (:1325-1350 + :377-445 + :1388-1410)
  const int mc = in.max_cars;
  int nc = in.n_cars[f];
  if (nc > mc) nc = mc;
  const int tl_in = in.target_lane_in[f];
  LaneStats ls;
  lane_stats_init(ls, cfg);
  Cand own, tl0, tl1, tl2;
  cand_init(own);
  cand_init(tl0);
  cand_init(tl1);
  cand_init(tl2);
  const double behind = em.s - cfg.car_length - cfg.safety_distance;  // :1402
  const int64_t cb = f * mc;
  for (int j = 0; j < nc; j++) {
    const int id = in.car_id[cb + j];
    const double x = in.car_x[cb + j], y = in.car_y[cb + j];
    const double vx = in.car_vx[cb + j], vy = in.car_vy[cb + j];
    const Match cm = lane_match(m, rs, x, y);
    double vs = 0, vd = 0;
    if (cm.ok) project_speed(m, vx, vy, cm.wp, vs, vd);
    if (out.car_lane) out.car_lane[cb + j] = cm.ok ? cm.lane : -1;
    if (out.car_next_wp) out.car_next_wp[cb + j] = cm.ok ? cm.wp : 0;
    if (out.car_s) out.car_s[cb + j] = cm.s;
    if (out.car_d) out.car_d[cb + j] = cm.d;
    if (out.car_vs) out.car_vs[cb + j] = vs;
    if (out.car_vd) out.car_vd[cb + j] = vd;
    if (!cm.ok) {  // :1336-1340 dropped from the map
      flags |= PP_F_CAR_DROPPED;
      continue;
    }
    lane_stats_add(ls, cfg, id, cm.lane, cm.s, vs, em.lane, tl_in, em.s, evs, e.dt0, flags);
    const double s0 = cm.s + vs * e.dt0;
    const double d0 = cm.d + vd * e.dt0;
    if (s0 > em.s && fabs(d0 - em.d) < 3) cand_offer(own, s0, id, j);
    if (s0 >= behind) {
      if (fabs(d0 - lane_center_offset(0)) < 3) cand_offer(tl0, s0, id, j);
      if (fabs(d0 - lane_center_offset(1)) < 3) cand_offer(tl1, s0, id, j);
      if (fabs(d0 - lane_center_offset(2)) < 3) cand_offer(tl2, s0, id, j);
    }
  }

  // ---- lane decision (:1355) + veto (:1358-1369)
  int target_lane = lane_stats_decide(ls, cfg, em.lane, tl_in);
  if (target_lane != em.lane) {
    const double dtl = lane_center_offset(target_lane);
    const double diff = fabs(evd * 1.0 + em.d - dtl);
    if (diff > 6.0) {
      flags |= PP_F_VETO;
      target_lane = em.lane;
    }
  }
  Cand tl = target_lane == 0 ? tl0 : (target_lane == 1 ? tl1 : tl2);
  if (tl.id == own.id) tl.id = -1;  // :1411 only check once

  // ---- speed target (:1422-1438)
  SpeedCtl sc;
  sc_init(sc, cfg, e.speed);
  if (own.id != -1) {
    double ts, tt;
    limit_speed(cfg, in.car_vx[cb + own.j], in.car_vy[cb + own.j], own.s0, em.s, e.speed, e.acc,
                true, ts, tt, flags);
    sc_limit(sc, ts, tt);
  }
  if (tl.id != -1) {
    double ts, tt;
    limit_speed(cfg, in.car_vx[cb + tl.j], in.car_vy[cb + tl.j], tl.s0, em.s, e.speed, e.acc,
                false, ts, tt, flags);
    sc_limit(sc, ts, tt);
  }
  if (out.target_speed) out.target_speed[f] = sc.target;
  if (out.target_time) out.target_time[f] = sc.time;

  // ---- trajectory (:1446-1448)
  const int np = build_trajectory(m, cfg, rs, in.prev_x + f * PP_PREV_KEEP,
                                  in.prev_y + f * PP_PREV_KEEP, e.nprev, e.x, e.y,
                                  in.ego_yaw_deg[f], target_lane, em.d, evd, sc,
                                  out.next_x + f * PP_PATH_LEN, out.next_y + f * PP_PATH_LEN, flags);

  for (int i = np; i < PP_PATH_LEN; i++) {  // short (fallback) paths: pad with NaN
    out.next_x[f * PP_PATH_LEN + i] = __longlong_as_double(0x7ff8000000000000ll);
    out.next_y[f * PP_PATH_LEN + i] = __longlong_as_double(0x7ff8000000000000ll);
  }
  out.n_points[f] = np;
  out.ego_lane[f] = em.lane;
  out.ref_wp[f] = rs.wp;
  out.target_lane[f] = target_lane;
  out.flags[f] = flags;
  if (out.ego_s) out.ego_s[f] = em.s;
  if (out.ego_d) out.ego_d[f] = em.d;
  if (out.ego_vs) out.ego_vs[f] = evs;
  if (out.ego_vd) out.ego_vd[f] = evd;
  if (out.ego_speed) out.ego_speed[f] = e.speed;
  if (out.ego_acc) out.ego_acc[f] = e.acc;
  if (out.next_car_id) out.next_car_id[f] = own.id;
  if (out.next_car_in_target_lane) out.next_car_in_target_lane[f] = tl.id;
}

// Throughput mapping: one thread (lane) per frame, persistent blocks.
__global__ void __launch_bounds__(kBlock)
plan_thread_per_frame(const double *__restrict__ map_table, int n_wp,
                      const __grid_constant__ pp_config cfg, const __grid_constant__ pp_frames in,
                      const __grid_constant__ pp_plans out, int64_t n_frames) {
  extern __shared__ double s_map[];
  const int words = n_wp * PP_MAP_STRIDE;
  for (int i = threadIdx.x; i < words; i += blockDim.x) s_map[i] = map_table[i];
  __syncthreads();
  MapView m{s_map, n_wp};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += stride)
    plan_frame(m, cfg, in, out, f);
}

// ---- aggregate statistics: exact int64 sums -----------------
__global__ void __launch_bounds__(256)
stats_kernel(const __grid_constant__ pp_plans p, int64_t n, unsigned long long *stats) {
  __shared__ unsigned long long s_acc[PP_STATS_LEN];
  for (int i = threadIdx.x; i < PP_STATS_LEN; i += blockDim.x) s_acc[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += stride) {
    const int np = p.n_points[f];
    const int tl = p.target_lane[f], el = p.ego_lane[f];
    const uint32_t fl = p.flags[f];
    atomicAdd(&s_acc[PP_STAT_FRAMES], 1ull);
    atomicAdd(&s_acc[PP_STAT_POINTS], (unsigned long long)np);
    if (tl >= 0 && tl < 3) atomicAdd(&s_acc[PP_STAT_TARGET_LANE0 + tl], 1ull);
    if (el >= 0 && el < 3) atomicAdd(&s_acc[PP_STAT_EGO_LANE0 + el], 1ull);
    if (tl != el) atomicAdd(&s_acc[PP_STAT_LANE_CHANGES], 1ull);
    for (int b = 0; b < PP_NUM_FLAGS; b++)
      if (fl & (1u << b)) atomicAdd(&s_acc[PP_STAT_FLAG0 + b], 1ull);
    long long xs = 0;  // fixed-point (1/256 m) checksum: order-independent
    for (int i = 0; i < np; i++) {
      const double x = p.next_x[f * PP_PATH_LEN + i], y = p.next_y[f * PP_PATH_LEN + i];
      if (x == x && y == y && fabs(x) < 1e12 && fabs(y) < 1e12)
        xs += (long long)(x * 256.0) + (long long)(y * 256.0);
    }
    atomicAdd(&s_acc[PP_STAT_XSUM], (unsigned long long)xs);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PP_STATS_LEN; i += blockDim.x)
    if (s_acc[i]) atomicAdd(&stats[i], s_acc[i]);
}

int g_variant = 0;
int g_sm_count = 0;

int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    g_sm_count = n;
  }
  return g_sm_count;
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ppi::set_cuda_error(what, (int)e, cudaGetErrorString(e));
    return PP_E_CUDA;
  }
  return PP_OK;
}

}  // namespace

extern "C" int pp_set_kernel_variant(int variant) {
  if (variant < 0 || variant > 2) return PP_E_ARG;
  g_variant = variant;
  return PP_OK;
}

extern "C" int pp_plan_batch(const pp_map *map, const pp_config *cfg, const pp_frames *in,
                             const pp_plans *out, int64_t n_frames, void *cuda_stream) {
  if (!map || !cfg || !in || !out || n_frames < 0) return PP_E_ARG;
  if (!map->dev_table) {
    ppi::set_cuda_error("pp_plan_batch: map has no device table (no usable CUDA device)", 0, "");
    return PP_E_CUDA;  // there is no CPU planning path
  }
  if (in->max_cars < 0 || in->max_cars > PP_MAX_CARS) return PP_E_RANGE;
  if (!in->ego_x || !in->ego_y || !in->ego_yaw_deg || !in->ego_speed_mph || !in->prev_n ||
      !in->prev_x || !in->prev_y || !in->target_lane_in || !in->n_cars)
    return PP_E_ARG;
  if (in->max_cars > 0 && (!in->car_id || !in->car_x || !in->car_y || !in->car_vx || !in->car_vy))
    return PP_E_ARG;
  if (!out->next_x || !out->next_y || !out->n_points || !out->ego_lane || !out->ref_wp ||
      !out->target_lane || !out->flags)
    return PP_E_ARG;
  if (n_frames == 0) return PP_OK;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const size_t smem = (size_t)map->n * PP_MAP_STRIDE * sizeof(double);
  if (smem > 200 * 1024) return PP_E_RANGE;
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    cudaFuncSetAttribute(plan_thread_per_frame, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    attr_set = true;
  }
  // persistent grid: a whole number of waves over the SMs, capped by the batch
  const int sms = sm_count();
  int64_t want = (n_frames + kBlock - 1) / kBlock;
  int64_t cap = (int64_t)sms * 8;
  int grid = (int)(want < cap ? want : cap);
  plan_thread_per_frame<<<grid, kBlock, smem, st>>>(map->dev_table, map->n, *cfg, *in, *out,
                                                    n_frames);
  ppi::count_launch();
  return check_launch("plan_thread_per_frame");
}

extern "C" int pp_stats_batch(const pp_plans *p, int64_t n_frames, int64_t *stats_dev,
                              void *cuda_stream) {
  if (!p || !stats_dev || n_frames < 0) return PP_E_ARG;
  if (!p->next_x || !p->next_y || !p->n_points || !p->ego_lane || !p->target_lane || !p->flags)
    return PP_E_ARG;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (cudaMemsetAsync(stats_dev, 0, PP_STATS_LEN * sizeof(int64_t), st) != cudaSuccess)
    return check_launch("pp_stats_batch memset");
  if (n_frames == 0) return PP_OK;
  int64_t want = (n_frames + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 4;
  int grid = (int)(want < cap ? want : cap);
  stats_kernel<<<grid, 256, 0, st>>>(*p, n_frames, (unsigned long long *)stats_dev);
  ppi::count_launch();
  return check_launch("stats_kernel");
}
