// pp_rollout.cu — closed-loop rollouts (BASELINE config 3): a device-resident
// simulator model around pp_plan_batch.  The model is specified in include/pp.h
// (pp_rollouts); every operation in it is + - * / (and one sqrt) so that the CPU
// restatement the tests check it against is bit-identical.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pp_internal.h"

struct pp_rollouts {
  const pp_map *map = nullptr;
  int64_t n = 0;   // rollouts
  int32_t c = 0;   // cars per rollout
  uint64_t seed = 0;
  int64_t first = 0;
  int64_t tick = 0;
  char *buf = nullptr;  // one device allocation
  // simulator state
  double *ego_x, *ego_y, *ego_yaw, *ego_mph, *car_ratio, *car_speed;
  int32_t *path_n, *path_off, *target_lane, *car_lane, *car_wp;
  bool lean = false;  // plans without the per-frame diagnostics and unused per-car outputs
  pp_plans full_pl;   // the complete set of output pointers (restored when lean is switched off)
  // per-tick frames and plans
  pp_frames fr;
  pp_plans pl;
  int64_t *stats_sum;  // running sum of the per-tick statistics over all ticks
  // Rollouts are independent, so they are cut into kGroups ranges that tick on their own
  // streams: one group's short kernels and side-stream tail overlap the other groups' work.
  static constexpr int kGroups = 8;  // streams that exist; PP_ROLLOUT_GROUPS of them are used
  cudaStream_t gs[kGroups] = {};
  cudaEvent_t g_done[kGroups] = {};
  cudaEvent_t fork = nullptr;
  cudaStream_t origin = nullptr;  // graph capture / replay stream (the caller's may be the
                                  // legacy default stream, which cannot be captured)
  cudaEvent_t o_fork = nullptr, o_join = nullptr;
  int64_t *tick_dev = nullptr;  // [R] ticks done per rollout (device side of `tick`)
  char *scratch[kGroups] = {};  // the planning pipeline's scratch, one per group
  // one tick of every group, captured once and replayed (a tick is ~50 small launches)
  cudaGraphExec_t graph = nullptr;
  int32_t graph_k = 0;
  pp_config graph_cfg;
  int64_t launches_per_tick = 0;
  int groups_wanted = 0;  // pp_rollouts_set_groups: 0 = automatic
};

namespace {

constexpr int kB = 256;
constexpr double kTick = 0.02;

// splitmix64, as in pp_synth.cpp
__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline double u01(uint64_t v) {
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// The map as the simulator sees it: rows of PP_MAP_STRIDE doubles, index wrapped into [0, n).
struct Track {
  const double *t;  // row 0
  int n;
  int stride;
  __host__ __device__ const double *row(int i) const {
    int k = i;
    if ((unsigned)i >= (unsigned)n) {  // the simulator keeps indices in [0, n): rarely taken
      k = i % n;
      if (k < 0) k += n;
    }
    return t + (size_t)k * stride;
  }
  __host__ __device__ double len(int w, int lane) const { return row(w)[10 + lane]; }
  // walk ds metres along lane `lane` from (w, u)
  __host__ __device__ void walk(int &w, double &u, int lane, double ds) const {
    for (int guard = 0; guard < 4 * n; guard++) {
      const double l = len(w, lane);
      if (ds >= 0) {
        const double rem = l * (1 - u);
        if (ds <= rem) {
          u += ds / l;
          break;
        }
        ds -= rem;
        u = 0;
        w++;
      } else {
        const double rem = l * u;
        if (-ds <= rem) {
          u += ds / l;
          break;
        }
        ds += rem;
        u = 1;
        w--;
      }
    }
    w %= n;
    if (w < 0) w += n;
  }
  // position and velocity of a car at (lane, w, u) moving at v
  __host__ __device__ void car(int lane, int w, double u, double v, double &x, double &y,
                               double &vx, double &vy) const {
    const double *a = row(w - 1), *b = row(w);
    const double ax = a[2 + 2 * lane], ay = a[3 + 2 * lane];
    const double bx = b[2 + 2 * lane], by = b[3 + 2 * lane];
    const double l = b[10 + lane];
    x = ax + (bx - ax) * u;
    y = ay + (by - ay) * u;
    const double tx = (bx - ax) / l, ty = (by - ay) / l;
    vx = v * tx;
    vy = v * ty;
  }
};

// Simulator state on the device (per rollout unless noted).  The unconsumed points of the last
// plan are not copied anywhere: they stay in the plan buffer and `path_off` says where they start.
struct SimState {
  double *ego_x, *ego_y, *ego_yaw, *ego_mph;
  int32_t *path_n, *path_off, *target_lane;
  int32_t *car_lane, *car_wp;       // [R][C]
  double *car_ratio, *car_speed;    // [R][C]
  int64_t *ticks;
};

// Both simulator kernels give a block of kB threads only kR rollouts: at 8,192 rollouts per stream
// group a tick is a chain of short dependent kernels, the job is bound by how long ONE frame
// takes to get through that chain (frames in flight / latency), and with a rollout per thread
// these two kernels were 31 and 68 us of it — twelve dependent trips over a thread's car slots,
// fifty over its trajectory row.  Eight threads per rollout make that two trips and seven.
constexpr int kR = 32;  // rollouts per block (one warp's worth: the ego part is one warp)

// frame <- state.  The block's threads sweep the rollouts' scalars (a warp per field), their
// kept points and their car slots, all coalesced.
__global__ void __launch_bounds__(kB, 6)
k_sim_frames(Track trk, int64_t lo, int64_t n, int c, SimState st, pp_plans pl, pp_frames fr) {
  const int64_t r0 = lo + (int64_t)blockIdx.x * kR;
  int64_t r_end = r0 + kR;
  if (r_end > lo + n) r_end = lo + n;
  const int nr = (int)(r_end - r0);
  {  // ego scalars: warp w copies field w
    const int field = threadIdx.x >> 5, rr = threadIdx.x & 31;
    const int64_t r = r0 + rr;
    if (rr < nr) {
      switch (field) {
        case 0: const_cast<double *>(fr.ego_x)[r] = st.ego_x[r]; break;
        case 1: const_cast<double *>(fr.ego_y)[r] = st.ego_y[r]; break;
        case 2: const_cast<double *>(fr.ego_yaw_deg)[r] = st.ego_yaw[r]; break;
        case 3: const_cast<double *>(fr.ego_speed_mph)[r] = st.ego_mph[r]; break;
        case 4: const_cast<int32_t *>(fr.prev_n)[r] = st.path_n[r]; break;
        case 5: const_cast<int32_t *>(fr.target_lane_in)[r] = st.target_lane[r]; break;
        case 6: const_cast<int32_t *>(fr.n_cars)[r] = c; break;
        default: break;
      }
    }
  }
  // kept points: the unconsumed points of the last plan, behind path_off
  for (int e = threadIdx.x; e < nr * PP_PREV_KEEP; e += kB) {
    const int rr = e / PP_PREV_KEEP, i = e - rr * PP_PREV_KEEP;
    const int64_t r = r0 + rr;
    const bool have = i < st.path_n[r];
    const int64_t src = r * PP_PATH_LEN + st.path_off[r] + i;
    const_cast<double *>(fr.prev_x)[r0 * PP_PREV_KEEP + e] = have ? pl.next_x[src] : 0.0;
    const_cast<double *>(fr.prev_y)[r0 * PP_PREV_KEEP + e] = have ? pl.next_y[src] : 0.0;
  }
  for (int64_t q = r0 * c + threadIdx.x; q < r_end * c; q += kB) {
    double x, y, vx, vy;
    trk.car(st.car_lane[q], st.car_wp[q], st.car_ratio[q], st.car_speed[q], x, y, vx, vy);
    const_cast<int32_t *>(fr.car_id)[q] = (int32_t)(q % c);
    const_cast<double *>(fr.car_x)[q] = x;
    const_cast<double *>(fr.car_y)[q] = y;
    const_cast<double *>(fr.car_vx)[q] = vx;
    const_cast<double *>(fr.car_vy)[q] = vy;
  }
}
static_assert(kB >= 7 * 32 && kR == 32, "k_sim_frames: a warp per scalar field, a lane per rollout");

// state <- simulator step(plan), and this tick's contribution to the aggregate statistics
// (definition: pp_stats_batch).  Warp 0 advances the egos, a lane per rollout, while the other
// warps sweep the car slots: the two parts touch different state except for the rollout's tick
// counter, which the respawn of a car reads, so only its increment waits behind the barrier
// (with the ego part after the barrier the block's other seven warps sat in it for 59 % of the
// kernel, profiles/r2_k_sim_advance_ncu.txt).  The counters are taken with warp votes, no
// atomics in shared memory.  The trajectory checksum normally comes from the planning kernels,
// which add it up as they write the points (plan_batch_scratch, xsum_add): re-reading 800 bytes
// per rollout for it was two thirds of this kernel's traffic.  For the single-kernel paths of
// small jobs the block takes it here, over its rollouts' rows (one contiguous run).
__global__ void __launch_bounds__(kB)
k_sim_advance(Track trk, int64_t lo, int64_t n, int c, uint64_t seed, int64_t first, int consume_k,
              SimState st, pp_plans pl, unsigned long long *stats_sum, int do_xsum) {
  const int64_t r0 = lo + (int64_t)blockIdx.x * kR;
  int64_t r_end = r0 + kR;
  if (r_end > lo + n) r_end = lo + n;
  const int nr = (int)(r_end - r0);
  if (threadIdx.x < 32) {
    // ---- the ego consumes k points of the new trajectory (one lane per rollout)
    const int ln = threadIdx.x;
    const bool live = ln < nr;
    const int64_t r = r0 + (live ? ln : 0);
    const int np = live ? pl.n_points[r] : 0;
    int tl = -1, el = -1;
    uint32_t fl = 0;
    if (live) {
      const int k = consume_k < np ? consume_k : np;
      if (k > 0) {
        const double ox = st.ego_x[r], oy = st.ego_y[r];
        const double qx = pl.next_x[r * PP_PATH_LEN + k - 1], qy = pl.next_y[r * PP_PATH_LEN + k - 1];
        const double d = sqrt((qx - ox) * (qx - ox) + (qy - oy) * (qy - oy));
        st.ego_x[r] = qx;
        st.ego_y[r] = qy;
        st.ego_mph[r] = d / (kTick * k) * 2.237;
      }
      st.path_n[r] = np - k;
      st.path_off[r] = k;
      tl = pl.target_lane[r];
      el = pl.ego_lane[r];
      st.target_lane[r] = tl;
      fl = pl.flags[r];
    }
    // counters: lane i ends up holding entry i of the statistics vector
    const unsigned full = 0xffffffffu;
    unsigned long long mine = 0;
    unsigned long long pts = (unsigned long long)np;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pts += __shfl_xor_sync(full, pts, o);
    const unsigned n_live = __popc(__ballot_sync(full, live));
    if (ln == PP_STAT_FRAMES) mine = n_live;
    if (ln == PP_STAT_POINTS) mine = pts;
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const unsigned a = __popc(__ballot_sync(full, live && tl == q));
      const unsigned b = __popc(__ballot_sync(full, live && el == q));
      if (ln == PP_STAT_TARGET_LANE0 + q) mine = a;
      if (ln == PP_STAT_EGO_LANE0 + q) mine = b;
    }
    {
      const unsigned a = __popc(__ballot_sync(full, live && tl != el));
      if (ln == PP_STAT_LANE_CHANGES) mine = a;
    }
#pragma unroll
    for (int b = 0; b < PP_NUM_FLAGS; b++) {
      const unsigned a = __popc(__ballot_sync(full, (fl >> b) & 1u));
      if (ln == PP_STAT_FLAG0 + b) mine = a;
    }
    if (ln < PP_STATS_LEN && ln != PP_STAT_XSUM && mine) atomicAdd(&stats_sum[ln], mine);
  } else {
    // ---- traffic (one thread per car slot; they read the rollout's tick before it is counted)
    for (int64_t q = r0 * c + (threadIdx.x - 32); q < r_end * c; q += kB - 32) {
      const int64_t r = q / c;
      const int j = (int)(q - r * c);
      const int np = pl.n_points[r];
      const int k = consume_k < np ? consume_k : np;
      const double dt = kTick * (k > 0 ? k : 1);
      int lane = st.car_lane[q], w = st.car_wp[q];
      double u = st.car_ratio[q], v = st.car_speed[q];
      const int m_lane = pl.car_lane[q];
      const double m_s = pl.car_s[q];
      if (m_lane < 0 || m_s < -100.0 || m_s > 300.0) {  // respawn
        const int64_t tick = st.ticks[r];
        const uint64_t key = mix64(mix64(seed ^ 0x5157ull) ^ ((uint64_t)(first + r) * 0xD1B54A32D192ED03ull));
        const uint64_t h = mix64(key ^ ((uint64_t)tick * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)j << 48));
        const double u1 = u01(mix64(h + 1)), u2 = u01(mix64(h + 2)), u3 = u01(mix64(h + 3));
        const double ds = (m_lane >= 0 && m_s < -100.0) ? 200.0 + 100.0 * u1 : -(60.0 + 40.0 * u1);
        lane = (int)(3.0 * u2);
        if (lane > 2) lane = 2;
        v = 17.88 + 8.94 * u3;
        w = pl.ref_wp[r];
        u = 0.5;
        trk.walk(w, u, lane, ds);
      } else {  // constant speed along the lane centre line
        u += (v * dt) / trk.len(w, lane);
        for (int guard = 0; guard < 64 && u >= 1; guard++) {
          const double left = (u - 1) * trk.len(w, lane);
          w = w + 1 == trk.n ? 0 : w + 1;
          u = left / trk.len(w, lane);
        }
      }
      st.car_lane[q] = lane;
      st.car_wp[q] = w;
      st.car_ratio[q] = u;
      st.car_speed[q] = v;
    }
  }
  if (do_xsum) {  // ---- checksum of the block's rows, unless the planning kernels added it
    long long xs = 0;
    const int64_t base = r0 * PP_PATH_LEN;
    for (int e = threadIdx.x; e < nr * PP_PATH_LEN; e += kB) {
      const int rr = e / PP_PATH_LEN;
      const int np_r = pl.n_points[r0 + rr];
      const double x = pl.next_x[base + e], y = pl.next_y[base + e];
      if (e - rr * PP_PATH_LEN < np_r && x == x && y == y && fabs(x) < 1e12 && fabs(y) < 1e12)
        xs += (long long)(x * 256.0) + (long long)(y * 256.0);
    }
    unsigned long long v = (unsigned long long)xs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&stats_sum[PP_STAT_XSUM], v);
  }
  __syncthreads();
  if (threadIdx.x < nr) st.ticks[r0 + threadIdx.x] += 1;
}
static_assert(PP_STATS_LEN <= 32, "k_sim_advance: a lane per statistics entry");

inline size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

int cuda_fail(const char *what, cudaError_t e) {
  ppi::set_cuda_error(what, (int)e, cudaGetErrorString(e));
  cudaGetLastError();
  return PP_E_CUDA;
}

}  // namespace

extern "C" int pp_rollouts_create(const pp_map *map, int64_t n, int32_t c, uint64_t seed,
                                  int64_t first, pp_rollouts **out) {
  if (!out) return PP_E_ARG;
  *out = nullptr;
  if (!map || n <= 0 || c < 0 || c > PP_MAX_CARS) return PP_E_ARG;
  {
    const int rc = ppi::check_map_device(map, "pp_rollouts_create");
    if (rc != PP_OK) return rc;
  }
  pp_rollouts *r = new (std::nothrow) pp_rollouts();
  if (!r) return PP_E_NOMEM;
  r->map = map;
  r->n = n;
  r->c = c;
  r->seed = seed;
  r->first = first;
  const size_t N = (size_t)n, NC = (size_t)n * (size_t)(c > 0 ? c : 1);
  // carve one allocation
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += al(bytes);
    return o;
  };
  const size_t o_ex = take(N * 8), o_ey = take(N * 8), o_yaw = take(N * 8), o_mph = take(N * 8);
  const size_t o_pn = take(N * 4), o_po = take(N * 4);
  const size_t o_tl = take(N * 4), o_cl = take(NC * 4), o_cw = take(NC * 4), o_cr = take(NC * 8),
               o_cs = take(NC * 8);
  // frames
  const size_t f_ex = take(N * 8), f_ey = take(N * 8), f_yaw = take(N * 8), f_mph = take(N * 8),
               f_pn = take(N * 4), f_px = take(N * PP_PREV_KEEP * 8), f_py = take(N * PP_PREV_KEEP * 8),
               f_tl = take(N * 4), f_nc = take(N * 4), f_id = take(NC * 4), f_cx = take(NC * 8),
               f_cy = take(NC * 8), f_vx = take(NC * 8), f_vy = take(NC * 8);
  // plans (all outputs: tests compare them tick by tick)
  const size_t p_nx = take(N * PP_PATH_LEN * 8), p_ny = take(N * PP_PATH_LEN * 8), p_np = take(N * 4),
               p_el = take(N * 4), p_rw = take(N * 4), p_tl = take(N * 4), p_fl = take(N * 4);
  size_t p_d[8];
  for (int i = 0; i < 8; i++) p_d[i] = take(N * 8);
  const size_t p_i0 = take(N * 4), p_i1 = take(N * 4);
  const size_t p_cs = take(NC * 8), p_cd = take(NC * 8), p_cvs = take(NC * 8), p_cvd = take(NC * 8),
               p_cl = take(NC * 4), p_cw = take(NC * 4);
  const size_t o_ss = take(PP_STATS_LEN * 8);
  const size_t o_tk = take(N * 8);
  cudaError_t e = cudaMalloc((void **)&r->buf, off);
  if (e != cudaSuccess) {
    delete r;
    return cuda_fail("cudaMalloc(rollouts)", e);
  }
  cudaMemset(r->buf, 0, off);
  char *b = r->buf;
  r->ego_x = (double *)(b + o_ex);
  r->ego_y = (double *)(b + o_ey);
  r->ego_yaw = (double *)(b + o_yaw);
  r->ego_mph = (double *)(b + o_mph);
  r->path_n = (int32_t *)(b + o_pn);
  r->path_off = (int32_t *)(b + o_po);
  r->target_lane = (int32_t *)(b + o_tl);
  r->car_lane = (int32_t *)(b + o_cl);
  r->car_wp = (int32_t *)(b + o_cw);
  r->car_ratio = (double *)(b + o_cr);
  r->car_speed = (double *)(b + o_cs);
  pp_frames &f = r->fr;
  f.ego_x = (double *)(b + f_ex);
  f.ego_y = (double *)(b + f_ey);
  f.ego_yaw_deg = (double *)(b + f_yaw);
  f.ego_speed_mph = (double *)(b + f_mph);
  f.prev_n = (int32_t *)(b + f_pn);
  f.prev_x = (double *)(b + f_px);
  f.prev_y = (double *)(b + f_py);
  f.target_lane_in = (int32_t *)(b + f_tl);
  f.n_cars = (int32_t *)(b + f_nc);
  f.car_id = (int32_t *)(b + f_id);
  f.car_x = (double *)(b + f_cx);
  f.car_y = (double *)(b + f_cy);
  f.car_vx = (double *)(b + f_vx);
  f.car_vy = (double *)(b + f_vy);
  f.max_cars = c > 0 ? c : 1;
  f.reserved = 0;
  pp_plans &p = r->pl;
  p.next_x = (double *)(b + p_nx);
  p.next_y = (double *)(b + p_ny);
  p.n_points = (int32_t *)(b + p_np);
  p.ego_lane = (int32_t *)(b + p_el);
  p.ref_wp = (int32_t *)(b + p_rw);
  p.target_lane = (int32_t *)(b + p_tl);
  p.flags = (uint32_t *)(b + p_fl);
  p.ego_s = (double *)(b + p_d[0]);
  p.ego_d = (double *)(b + p_d[1]);
  p.ego_vs = (double *)(b + p_d[2]);
  p.ego_vd = (double *)(b + p_d[3]);
  p.ego_speed = (double *)(b + p_d[4]);
  p.ego_acc = (double *)(b + p_d[5]);
  p.target_speed = (double *)(b + p_d[6]);
  p.target_time = (double *)(b + p_d[7]);
  p.next_car_id = (int32_t *)(b + p_i0);
  p.next_car_in_target_lane = (int32_t *)(b + p_i1);
  p.car_s = (double *)(b + p_cs);
  p.car_d = (double *)(b + p_cd);
  p.car_vs = (double *)(b + p_cvs);
  p.car_vd = (double *)(b + p_cvd);
  p.car_lane = (int32_t *)(b + p_cl);
  p.car_next_wp = (int32_t *)(b + p_cw);
  r->stats_sum = (int64_t *)(b + o_ss);
  r->tick_dev = (int64_t *)(b + o_tk);

  // ---- initial state on the host: a pure function of (seed, first + r)
  std::vector<double> ex(N), ey(N), yaw(N), mph(N), cr(NC), cs(NC);
  std::vector<int32_t> tl(N), cl(NC), cw(NC);
  Track trk{map->table.data(), map->n, PP_MAP_STRIDE};
  for (int64_t i = 0; i < n; i++) {
    uint64_t key = mix64(mix64(seed ^ 0x1417ull) ^ ((uint64_t)(first + i) * 0xD1B54A32D192ED03ull));
    uint64_t ctr = 0;
    auto uni = [&]() { return u01(mix64(key + (ctr++) * 0x9E3779B97F4A7C15ull)); };
    int w = (int)(uni() * map->n);
    if (w >= map->n) w = map->n - 1;
    double u = uni();
    int lane = (int)(uni() * 3);
    if (lane > 2) lane = 2;
    double x, y, vx, vy;
    trk.car(lane, w, u, 1.0, x, y, vx, vy);
    ex[i] = x;
    ey[i] = y;
    yaw[i] = std::atan2(vy, vx) * 180 / M_PI;
    mph[i] = 0.0;
    tl[i] = lane;
    for (int j = 0; j < c; j++) {
      int l = (int)(uni() * 3);
      if (l > 2) l = 2;
      const double ds = -100.0 + 400.0 * uni();
      const double sp = 17.88 + 8.94 * uni();
      int ww = w;
      double uu = u;
      trk.walk(ww, uu, l, ds);
      cl[i * c + j] = l;
      cw[i * c + j] = ww;
      cr[i * c + j] = uu;
      cs[i * c + j] = sp;
    }
  }
  bool ok = true;
  auto up = [&](void *dst, const void *src, size_t bytes) {
    if (bytes && cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
  };
  up(r->ego_x, ex.data(), N * 8);
  up(r->ego_y, ey.data(), N * 8);
  up(r->ego_yaw, yaw.data(), N * 8);
  up(r->ego_mph, mph.data(), N * 8);
  up(r->target_lane, tl.data(), N * 4);
  if (c > 0) {
    up(r->car_lane, cl.data(), NC * 4);
    up(r->car_wp, cw.data(), NC * 4);
    up(r->car_ratio, cr.data(), NC * 8);
    up(r->car_speed, cs.data(), NC * 8);
  }
  if (!ok) {
    cudaError_t e2 = cudaGetLastError();
    cudaFree(r->buf);
    delete r;
    return cuda_fail("cudaMemcpy(rollouts init)", e2);
  }
  for (int g = 0; g < pp_rollouts::kGroups; g++) {
    if (cudaStreamCreateWithFlags(&r->gs[g], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->g_done[g], cudaEventDisableTiming) != cudaSuccess)
      ok = false;
  }
  if (cudaEventCreateWithFlags(&r->fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&r->o_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&r->o_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&r->origin, cudaStreamNonBlocking) != cudaSuccess)
    ok = false;
  if (!ok) {
    cudaError_t e2 = cudaGetLastError();
    pp_rollouts_destroy(r);
    return cuda_fail("rollout streams", e2);
  }
  *out = r;
  return PP_OK;
}

extern "C" void pp_rollouts_destroy(pp_rollouts *r) {
  if (!r) return;
  for (int g = 0; g < pp_rollouts::kGroups; g++) {
    if (r->scratch[g]) cudaFree(r->scratch[g]);
    if (r->gs[g]) cudaStreamDestroy(r->gs[g]);
    if (r->g_done[g]) cudaEventDestroy(r->g_done[g]);
  }
  if (r->fork) cudaEventDestroy(r->fork);
  if (r->graph) cudaGraphExecDestroy(r->graph);
  if (r->origin) cudaStreamDestroy(r->origin);
  if (r->o_fork) cudaEventDestroy(r->o_fork);
  if (r->o_join) cudaEventDestroy(r->o_join);
  if (r->buf) cudaFree(r->buf);
  delete r;
}

namespace {

// One tick of every group: issued on the group streams, forked from / joined to `st`.
// One tick of every group.  Groups are independent (a rollout's tick t+1 depends on its own tick
// t only), so a run forks the group streams off `st` once (`fork`) and joins them once (`join`);
// a captured tick needs both so that the graph is closed.
int issue_tick(pp_rollouts *r, const pp_config *cfg, int32_t consume_k, cudaStream_t st,
               bool fork = true, bool join = true) {
  // device rows of the padded table: logical row 0 starts PPD_PAD_ROWS rows in
  const Track trk{r->map->dev_table + (size_t)PPD_PAD_ROWS * PP_MAP_STRIDE, r->map->n, PP_MAP_STRIDE};
  // groups: contiguous ranges of rollouts, at least 16,384 each, at most 8.  A tick of one group
  // is a chain of eight dependent kernels whose durations simply add up (253 us at 65,536
  // rollouts under ncu, 254 us per tick measured, profiles/r2_rollouts_chain.csv); with several
  // groups different kernels of different groups overlap.  65,536 rollouts, four runs each
  // (profiles/r2_rollouts_groups2.log): 1 group 253-256 M ego-frames/s, 2 groups 271-272 M,
  // 4 groups 270-283 M, 8 groups 254-263 M (64 launches per tick; the host falls behind); any
  // of them loses a run now and then to the shared host (79-183 M).
  // pp_rollouts_set_groups / PP_ROLLOUT_GROUPS override the count (at least 4,096 each then).
  static const int forced_groups = [] {
    const char *e = getenv("PP_ROLLOUT_GROUPS");
    const int v = e && *e ? atoi(e) : 0;
    return v < 0 ? 0 : (v > pp_rollouts::kGroups ? pp_rollouts::kGroups : v);
  }();
  const int forced = r->groups_wanted ? r->groups_wanted : forced_groups;
  int groups = forced ? forced : pp_rollouts::kGroups;
  while (groups > 1 && r->n / groups < (forced ? 4096 : 16384)) groups--;
  const int64_t per = (r->n + groups - 1) / groups;
  const int mc = r->fr.max_cars;
  if (fork) {
    cudaEventRecord(r->fork, st);
    for (int g = 0; g < groups; g++) cudaStreamWaitEvent(r->gs[g], r->fork, 0);
  }
  const int64_t launches0 = pp_launch_count();
  for (int g = 0; g < groups; g++) {
    const int64_t lo = g * per;
    const int64_t cnt = (r->n - lo) < per ? (r->n - lo) : per;
    if (cnt <= 0) continue;
    cudaStream_t gs = r->gs[g];
    const int grid = (int)((cnt + kR - 1) / kR);
    const pp_frames fr = ppi::offset_frames(r->fr, lo);
    const pp_plans pl = ppi::offset_plans(r->pl, lo, mc);
    const SimState sst{r->ego_x, r->ego_y, r->ego_yaw, r->ego_mph, r->path_n, r->path_off,
                       r->target_lane, r->car_lane, r->car_wp, r->car_ratio, r->car_speed, r->tick_dev};
    k_sim_frames<<<grid, kB, 0, gs>>>(trk, lo, cnt, r->c, sst, r->pl, r->fr);
    if (!r->scratch[g]) {
      const size_t need = ppi::plan_scratch_bytes(per, mc);
      if (need && cudaMalloc((void **)&r->scratch[g], need) != cudaSuccess)
        return cuda_fail("cudaMalloc(rollout scratch)", cudaGetLastError());
    }
    const bool plan_sums = ppi::plan_adds_checksum(cnt);
    int rc = ppi::plan_batch_scratch(r->map, cfg, &fr, &pl, cnt, gs, r->scratch[g], nullptr,
                                     (unsigned long long *)r->stats_sum + PP_STAT_XSUM);
    if (rc != PP_OK) return rc;
    k_sim_advance<<<grid, kB, 0, gs>>>(trk, lo, cnt, r->c, r->seed, r->first, consume_k, sst, r->pl,
                                       (unsigned long long *)r->stats_sum, plan_sums ? 0 : 1);
    ppi::count_launch(2);
  }
  r->launches_per_tick = pp_launch_count() - launches0;
  if (join) {
    for (int g = 0; g < groups; g++) {
      cudaEventRecord(r->g_done[g], r->gs[g]);
      cudaStreamWaitEvent(st, r->g_done[g], 0);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("rollout tick", e);
  return PP_OK;
}

void issue_join_only(pp_rollouts *r, cudaStream_t st) {
  for (int g = 0; g < pp_rollouts::kGroups; g++) {
    cudaEventRecord(r->g_done[g], r->gs[g]);
    cudaStreamWaitEvent(st, r->g_done[g], 0);
  }
}

bool same_cfg(const pp_config &a, const pp_config &b) { return std::memcmp(&a, &b, sizeof a) == 0; }

}  // namespace

extern "C" int pp_rollouts_run(pp_rollouts *r, const pp_config *cfg, int64_t n_ticks,
                               int32_t consume_k, void *cuda_stream) {
  if (!r || !cfg || n_ticks < 0 || consume_k < 1 || consume_k > PP_PATH_LEN - PP_PREV_KEEP)
    return PP_E_ARG;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (n_ticks == 0) return PP_OK;
  {
    const int rc = ppi::check_map_device(r->map, "pp_rollouts_run");
    if (rc != PP_OK) return rc;
  }
  // A tick is ~36 short launches over 8 streams (4 groups, each with its side stream).  It can
  // be captured once and replayed as a CUDA graph (PP_ROLLOUT_GRAPH=1; the pipeline's scratch is
  // owned by the rollouts object, so the graph holds only kernels, memsets and event edges).
  // With enough hardware work queues (pp_api.cu) both ways run at the same, stable speed
  // (≈ 200 M ego-frames/s for 65,536 rollouts on one B200); direct issue is the default because
  // it needs no per-tick join of the groups.
  const bool use_graph = n_ticks >= 8 && getenv("PP_ROLLOUT_GRAPH") != nullptr;
  if (use_graph) {
    cudaStream_t caller = st;
    cudaEventRecord(r->o_fork, caller);
    st = r->origin;
    cudaStreamWaitEvent(st, r->o_fork, 0);
    if (r->graph && (r->graph_k != consume_k || !same_cfg(r->graph_cfg, *cfg))) {
      cudaGraphExecDestroy(r->graph);
      r->graph = nullptr;
    }
    if (!r->graph) {
      // one direct tick first: sizes the stream-ordered pool and sets kernel attributes
      int rc = issue_tick(r, cfg, consume_k, st);
      if (rc != PP_OK) return rc;
      r->tick++;
      n_ticks--;
      cudaGraph_t g = nullptr;
      cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
      if (e != cudaSuccess) return cuda_fail("cudaStreamBeginCapture", e);
      rc = issue_tick(r, cfg, consume_k, st);
      e = cudaStreamEndCapture(st, &g);
      if (rc != PP_OK || e != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        return rc != PP_OK ? rc : cuda_fail("cudaStreamEndCapture", e);
      }
      e = cudaGraphInstantiate(&r->graph, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) {
        r->graph = nullptr;
        return cuda_fail("cudaGraphInstantiate", e);
      }
      r->graph_k = consume_k;
      r->graph_cfg = *cfg;
    }
    for (int64_t t = 0; t < n_ticks; t++) {
      cudaError_t e = cudaGraphLaunch(r->graph, st);
      if (e != cudaSuccess) return cuda_fail("cudaGraphLaunch", e);
      ppi::count_launch((int)r->launches_per_tick);
      r->tick++;
    }
    cudaEventRecord(r->o_join, st);
    cudaStreamWaitEvent(caller, r->o_join, 0);
    return PP_OK;
  }
  for (int64_t t = 0; t < n_ticks; t++) {
    const int rc = issue_tick(r, cfg, consume_k, st, t == 0, t == n_ticks - 1);
    if (rc != PP_OK) {
      issue_join_only(r, st);  // keep `st` ordered after whatever the group streams were given
      return rc;
    }
    r->tick++;
  }
  return PP_OK;
}

extern "C" int pp_rollouts_set_lean(pp_rollouts *r, int lean) {
  if (!r) return PP_E_ARG;
  if (lean && !r->lean) {
    r->full_pl = r->pl;
    pp_plans &p = r->pl;
    p.ego_s = p.ego_d = p.ego_vs = p.ego_vd = p.ego_speed = p.ego_acc = nullptr;
    p.target_speed = p.target_time = nullptr;
    p.next_car_id = p.next_car_in_target_lane = nullptr;
    p.car_d = p.car_vs = p.car_vd = nullptr;
    p.car_next_wp = nullptr;
  } else if (!lean && r->lean) {
    r->pl = r->full_pl;
  }
  r->lean = lean != 0;
  if (r->graph) {  // the captured tick holds the old pointers
    cudaGraphExecDestroy(r->graph);
    r->graph = nullptr;
  }
  return PP_OK;
}

extern "C" int pp_rollouts_set_groups(pp_rollouts *r, int groups) {
  if (!r || groups < 0 || groups > pp_rollouts::kGroups) return PP_E_ARG;
  if (r->graph && groups != r->groups_wanted) {  // the captured tick holds the old split
    cudaGraphExecDestroy(r->graph);
    r->graph = nullptr;
  }
  r->groups_wanted = groups;
  return PP_OK;
}

extern "C" int pp_rollouts_last(const pp_rollouts *r, pp_frames *frames_dev, pp_plans *plans_dev) {
  if (!r) return PP_E_ARG;
  if (frames_dev) *frames_dev = r->fr;
  if (plans_dev) *plans_dev = r->pl;
  return PP_OK;
}

extern "C" int pp_rollouts_get_state(const pp_rollouts *r, pp_rollout_state *h) {
  if (!r || !h) return PP_E_ARG;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cuda_fail("pp_rollouts_get_state", e);
  const size_t N = (size_t)r->n, NC = (size_t)r->n * (size_t)r->c;
  bool ok = true;
  auto dn = [&](void *dst, const void *src, size_t bytes) {
    if (dst && bytes && cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
  };
  dn(h->ego_x, r->ego_x, N * 8);
  dn(h->ego_y, r->ego_y, N * 8);
  dn(h->ego_yaw_deg, r->ego_yaw, N * 8);
  dn(h->ego_speed_mph, r->ego_mph, N * 8);
  dn(h->path_n, r->path_n, N * 4);
  if (h->path_x || h->path_y) {  // the unconsumed points live in the plan buffer at path_off
    std::vector<int32_t> pn(N), po(N);
    std::vector<double> buf(N * PP_PATH_LEN);
    dn(pn.data(), r->path_n, N * 4);
    dn(po.data(), r->path_off, N * 4);
    for (int axis = 0; axis < 2; axis++) {
      double *dst = axis ? h->path_y : h->path_x;
      if (!dst) continue;
      dn(buf.data(), axis ? r->pl.next_y : r->pl.next_x, N * PP_PATH_LEN * 8);
      for (size_t i = 0; i < N; i++)
        for (int k = 0; k < PP_PATH_LEN; k++)
          dst[i * PP_PATH_LEN + k] = k < pn[i] ? buf[i * PP_PATH_LEN + po[i] + k] : 0.0;
    }
  }
  dn(h->target_lane, r->target_lane, N * 4);
  dn(h->car_lane, r->car_lane, NC * 4);
  dn(h->car_wp, r->car_wp, NC * 4);
  dn(h->car_ratio, r->car_ratio, NC * 8);
  dn(h->car_speed, r->car_speed, NC * 8);
  h->tick = r->tick;
  if (!ok) return cuda_fail("pp_rollouts_get_state copy", cudaGetLastError());
  return PP_OK;
}

extern "C" int pp_rollouts_stats(const pp_rollouts *r, int64_t *stats_dev, void *cuda_stream) {
  if (!r || !stats_dev) return PP_E_ARG;
  cudaError_t e = cudaMemcpyAsync(stats_dev, r->stats_sum, PP_STATS_LEN * sizeof(int64_t),
                                  cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail("pp_rollouts_stats", e);
  return PP_OK;
}
