/* pp.h — C ABI of the B200-native batched highway path planner.
 *
 * One call, pp_plan_batch, replaces the planning step that the reference runs
 * once per telemetry message inside main::onMessage
 * (reference src/main.cpp:1254-1457): ego state from the 10th reused point,
 * Map::init_reference_waypoint / lane_matching / project_speed, the
 * sensor-fusion matching loop, LaneChangePlanner::calculate_target_lane,
 * LimitSpeed::calculate, SpeedController and TrajectoryBuilder::build
 * (tk::spline fit + 0.02 s point emission) — for N independent frames at once,
 * on the GPU, over struct-of-arrays frame buffers.
 *
 * The reference has no FFI of its own for this path (the path is inlined in a
 * lambda whose only interface is the simulator wire message), so every entry
 * point below cites the reference lines whose behaviour it reproduces.
 *
 * Conventions
 *   - plain C, no exceptions cross this boundary; every function returns
 *     PP_OK (0) or a negative PP_E_* code; pp_strerror() names it.
 *   - "dev" pointers are CUDA device pointers on the device that was current
 *     when the pp_map was created; "host" pointers are ordinary (ideally
 *     pinned) host memory.  The caller owns every buffer; the library owns
 *     pp_map.  A call made while another device is current fails with
 *     PP_E_ARG (one pp_map per device).
 *   - pp_plan_batch is asynchronous on the caller's stream and re-entrant:
 *     the reference keeps reference_waypoint_id/ratio as mutable state on
 *     the shared Map (src/main.cpp:132-133); here that state lives in
 *     registers per frame, so the map is immutable after creation.
 *   - There is NO CPU planning path in this library.  If the CUDA device or
 *     kernels are unavailable the calls fail with PP_E_CUDA.
 */
#ifndef PP_B200_H
#define PP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_VERSION 102

/* Fixed sizes of the reference's planning step. */
#define PP_NUM_LANES 3    /* src/main.cpp:22  NUM_LANES */
#define PP_PREV_KEEP 10   /* src/main.cpp:1258 prev_trajectory_length */
#define PP_PATH_LEN 50    /* src/main.cpp:854,1039 result_points.size() < 50 */
#define PP_MAX_CARS 64    /* largest n_cars per frame this build accepts (BASELINE config 5) */

/* One row of the uploaded map table (doubles per waypoint):
 * ref.x ref.y c0.x c0.y c1.x c1.y c2.x c2.y nx ny len0 len1 len2
 * = Map::Waypoint (src/main.cpp:76-82) plus get_lane_length(i, lane)
 * (src/main.cpp:138-142) precomputed with the same expression. */
#define PP_MAP_STRIDE 13

/* Error codes. */
#define PP_OK 0
#define PP_E_ARG (-1)      /* null pointer / bad size */
#define PP_E_CUDA (-2)     /* CUDA runtime error (pp_last_cuda_error() has the text) */
#define PP_E_IO (-3)       /* map file could not be read */
#define PP_E_NOMEM (-4)
#define PP_E_RANGE (-5)    /* n_cars > PP_MAX_CARS, n_waypoints out of range, ... */

/* Per-frame flag bits: one bit per print / log site of the reference, so
 * that anomalies are data and never abort the batch. */
#define PP_F_EGO_MATCH_FAIL   (1u << 0)  /* src/main.cpp:1304 "can't lane match ego" */
#define PP_F_CAR_DROPPED      (1u << 1)  /* :1338 car failed lane_matching, erased */
#define PP_F_COLLISION        (1u << 2)  /* :1077 "detected collision" */
#define PP_F_BRAKE            (1u << 3)  /* :1109 */
#define PP_F_MAXBRAKE         (1u << 4)  /* :1101 */
#define PP_F_ADJUST           (1u << 5)  /* :1131 */
#define PP_F_KEEP             (1u << 6)  /* :1146 */
#define PP_F_SPLINE_INPUT_ERR (1u << 7)  /* :837 non-increasing knot x, truncated */
#define PP_F_FALLBACK         (1u << 8)  /* :848 angle-based generator taken (silent in the reference) */
#define PP_F_ACC_OVERRIDE     (1u << 9)  /* :964 */
#define PP_F_CURV_ADJUST      (1u << 10) /* :988 */
#define PP_F_LANE_SWITCH_NEG  (1u << 11) /* :711 */
#define PP_F_VETO             (1u << 12) /* :1366 "target lane too far" */
#define PP_F_ACCT_HIGH        (1u << 13) /* :950 */
#define PP_F_ACCN_HIGH        (1u << 14) /* :978 */
#define PP_F_SPLINE_WARNING   (1u << 15) /* :927 */
#define PP_F_CLOSED_RANGE     (1u << 16) /* :408 condition true for some car */
#define PP_F_CLOSED_AHEAD     (1u << 17) /* :423 condition true for some car */
#define PP_F_CLOSED_BEHIND    (1u << 18) /* :439 condition true for some car */
#define PP_F_TRANSFORM_ERR    (1u << 19) /* :1015 */
#define PP_F_COLD_START       (1u << 20) /* :1261 fewer than 10 previous points */
#define PP_NUM_FLAGS 21

typedef struct pp_map pp_map; /* opaque: host table + device copy */

/* Tunables = the reference's globals (src/main.cpp:30,39-49).
 * pp_config_default() fills in exactly those literals. */
typedef struct pp_config {
  double relaxed_acc;                   /* 5    :39 */
  double min_relaxed_acc_while_braking; /* 4    :40 */
  double maximum_acc;                   /* 8    :42 */
  double max_speed;                     /* 22.2 :45 */
  double car_length;                    /* 4.5  :46 */
  double safety_distance;               /* 2    :47 */
  double keep_distance;                 /* 10   :48 */
  double keep_distance_leeway;          /* 0.5  :49 */
  int32_t test_fast_lane_change;        /* 0    :30 */
  int32_t reserved;
} pp_config;

/* Input frames, struct of arrays, N frames.  Inner arrays are frame-major:
 * prev_x[f*PP_PREV_KEEP + i], car_x[f*max_cars + j].
 * Fields mirror the telemetry the reference reads (src/main.cpp:1233-1252,
 * 1297,1328-1334); the unused wire fields (s, d, end_path_s/d, car s/d) are
 * not carried.  Car ids within one frame must be distinct (the reference's map keeps the
 * LAST row of a repeated id; pp::wire::parse_telemetry applies that rule); cars may appear
 * in any order (the reference iterates a std::map<int,Car>, i.e. ascending id, and only
 * tie-breaks depend on that order — reproduced here by id). */
typedef struct pp_frames {
  const double *ego_x;         /* [N] telemetry x  (used when prev_n < 10, :1233) */
  const double *ego_y;         /* [N] */
  const double *ego_yaw_deg;   /* [N] degrees; heading fallback only (:587,596,604) */
  const double *ego_speed_mph; /* [N] used when prev_n < 10 (:1238-1239) */
  const int32_t *prev_n;       /* [N] previous_path size; only ">= 10" matters (:1261) */
  const double *prev_x;        /* [N][10] first 10 points of previous_path_x */
  const double *prev_y;        /* [N][10] */
  const int32_t *target_lane_in; /* [N] persistent target_lane (:1195,1355) */
  const int32_t *n_cars;       /* [N] 0..max_cars */
  const int32_t *car_id;       /* [N][max_cars] */
  const double *car_x;         /* [N][max_cars] */
  const double *car_y;
  const double *car_vx;
  const double *car_vy;
  int32_t max_cars;            /* row length of the car arrays, <= PP_MAX_CARS */
  int32_t reserved;
  /* Optional, all five NULL or all five set: cars the reference still holds in its persistent
   * std::map<int,Car> although the current message does not list them (src/main.cpp:1194,
   * 1325-1334 only touch the cars of the message).  Such a car takes part in the planning with
   * the Frenet values of its last sighting, which the reference never recomputes:
   * car_frozen_lane[f*max_cars + j] >= 0 marks slot j as frozen (lane = that value, s / d / vs /
   * vd from the arrays below; its car_x / car_y are not matched, car_vx / car_vy still give its
   * Cartesian speed to LimitSpeed, :1080); -1 = an ordinary car of the current message.
   * include/pp_wire.hpp (pp::wire::Session) maintains this state for replayed sessions. */
  const int32_t *car_frozen_lane; /* [N][max_cars] */
  const double *car_frozen_s;     /* [N][max_cars] */
  const double *car_frozen_d;
  const double *car_frozen_vs;
  const double *car_frozen_vd;
} pp_frames;

/* Output plans, struct of arrays.  next_x/next_y entries at and beyond
 * n_points are set to quiet NaN.  Any pointer in the "diagnostics" and "per-car" groups may
 * be NULL (that output is then skipped). */
typedef struct pp_plans {
  double *next_x;        /* [N][50]  (:1450-1462) */
  double *next_y;        /* [N][50] */
  int32_t *n_points;     /* [N] <= 50 (10 kept + <= 40 new; fewer only in the fallback) */
  int32_t *ego_lane;     /* [N] */
  int32_t *ref_wp;       /* [N] reference_waypoint_id, un-wrapped, 0..n_wp (:186) */
  int32_t *target_lane;  /* [N] after planner + veto: the state carried to the next frame */
  uint32_t *flags;       /* [N] PP_F_* */
  /* diagnostics */
  double *ego_s, *ego_d, *ego_vs, *ego_vd;  /* [N] (:1302,1313) */
  double *ego_speed, *ego_acc;              /* [N] (:1273-1276,1319-1320) */
  double *target_speed, *target_time;       /* [N] SpeedController after both limits (:1430,1437) */
  int32_t *next_car_id;                     /* [N] -1 if none (:1383-1400) */
  int32_t *next_car_in_target_lane;         /* [N] -1 if none (:1402-1411) */
  /* per-car results of the sensor-fusion matching loop (:1325-1350) */
  double *car_s, *car_d, *car_vs, *car_vd;  /* [N][max_cars] */
  int32_t *car_lane;                        /* [N][max_cars]; -1 = dropped */
  int32_t *car_next_wp;                     /* [N][max_cars] un-wrapped */
} pp_plans;

/* Aggregate statistics vector (all int64 so that sums are exact and the
 * 1/2/4/8-GPU all-reduce is bit-identical). */
#define PP_STAT_FRAMES 0
#define PP_STAT_POINTS 1                         /* sum of n_points */
#define PP_STAT_TARGET_LANE0 2                   /* +lane */
#define PP_STAT_EGO_LANE0 5                      /* +lane */
#define PP_STAT_LANE_CHANGES 8                   /* target_lane != ego_lane */
#define PP_STAT_FLAG0 9                          /* +bit, PP_NUM_FLAGS entries */
#define PP_STAT_XSUM (PP_STAT_FLAG0 + PP_NUM_FLAGS) /* fixed-point checksum of emitted x,y */
#define PP_STATS_LEN (PP_STAT_XSUM + 1)

int pp_version(void);
/* Optional, before the process's first CUDA call: asks for 32 hardware work queues
 * (CUDA_DEVICE_MAX_CONNECTIONS, unless the environment already sets it).  The planner keeps up
 * to nine streams busy; with the default 8 queues independent streams share a queue and
 * serialise.  The library never changes the environment by itself. */
int pp_init(void);
const char *pp_strerror(int code);
const char *pp_last_cuda_error(void);
int pp_device_count(void); /* number of CUDA devices, or PP_E_CUDA */

/* Defaults = the literals at src/main.cpp:30,39-49. */
int pp_config_default(pp_config *cfg);

/* Map::Init (src/main.cpp:89-131) on the host (same libm as the reference, so
 * the table is bit-identical), then one upload to the current device. */
int pp_map_create(const double *wx, const double *wy, int n, pp_map **out);
/* CSV loader with the reference's parsing (x,y as double; columns 3-5 ignored;
 * src/main.cpp:1171-1191). */
int pp_map_create_from_csv(const char *path, pp_map **out);
void pp_map_destroy(pp_map *map);
int pp_map_num_waypoints(const pp_map *map);
/* Copy the host table out: n * PP_MAP_STRIDE doubles. */
int pp_map_table(const pp_map *map, double *out);

/* The hot path.  in/out hold DEVICE pointers.  Asynchronous on `cuda_stream`
 * (a cudaStream_t, may be NULL for the default stream). */
int pp_plan_batch(const pp_map *map, const pp_config *cfg, const pp_frames *in,
                  const pp_plans *out, int64_t n_frames, void *cuda_stream);

/* Same call with HOST buffers: uploads the inputs, plans, downloads the
 * outputs (chunked, copies overlapped with compute), returns when the host
 * buffers are complete.  This is the drop-in for a CPU caller. */
int pp_plan_batch_host(const pp_map *map, const pp_config *cfg, const pp_frames *in,
                       const pp_plans *out, int64_t n_frames);

/* The same call for a caller that does not want its own points sent back over PCIe.  A frame
 * with prev_n >= PP_PREV_KEEP keeps its first PP_PREV_KEEP previous points as the first points
 * of the new trajectory, verbatim (result_points = prev_trajectory, src/main.cpp:578,
 * 1446-1452) — the caller already holds them in in->prev_x / in->prev_y.  Here a trajectory
 * comes back in two parts, and out->next_x / out->next_y must be NULL:
 *   tail_x, tail_y [n][PP_PATH_LEN - PP_PREV_KEEP]  points PP_PREV_KEEP.. of every frame;
 *   head_x, head_y [n][PP_PREV_KEEP]                points 0..PP_PREV_KEEP-1.  The library writes
 *       a head row ONLY for a frame that kept nothing (prev_n < PP_PREV_KEEP: all its points are
 *       new); every other row is left as the caller has it.  head_x / head_y may be the very
 *       buffers passed as in->prev_x / in->prev_y (they are read before they are written), and
 *       then hold the first points of every trajectory on return.
 * head ++ tail of frame i equals row i of pp_plan_batch_host's next_x / next_y bit for bit
 * (tests/test_gpu_parity.py); 160 of the 820 bytes per frame stay off the bus. */
typedef struct pp_split_rows {
  double *head_x, *head_y;
  double *tail_x, *tail_y;
} pp_split_rows;
int pp_plan_batch_host_split(const pp_map *map, const pp_config *cfg, const pp_frames *in,
                             const pp_plans *out, const pp_split_rows *rows, int64_t n_frames);

/* Aggregate statistics of a planned batch on the device:
 * stats_dev[PP_STATS_LEN] (int64, overwritten).  The multi-GPU job all-reduces
 * this vector (pp_stats_reduce, ncclSum) — the only collective on the path. */
int pp_stats_batch(const pp_plans *plans_dev, int64_t n_frames, int64_t *stats_dev,
                   void *cuda_stream);

/* f64 minimum / maximum statistics of a planned batch, computed from the plan
 * outputs alone so that any caller can re-derive them: with V_k = (P_{k+1} - P_k) * 50 and
 * A_k = (V_{k+1} - V_k) * 50 over a frame's n_points output points (k < n_points - 1 and
 * k < n_points - 2), every operation in IEEE double, left to right, no fused multiply-add.
 * Non-finite values do not take part; an entry nothing contributed to stays at +inf (minima)
 * or -inf (maxima).  ego_speed / target_speed are optional outputs of pp_plans. */
#define PP_FSTAT_MIN_EGO_SPEED 0     /* min over frames of ego_speed (:1273) */
#define PP_FSTAT_MIN_TARGET_SPEED 1  /* ... of the SpeedController target (:1430,1437) */
#define PP_FSTAT_MIN_STEP_SPEED 2    /* min over frames and k of |V_k| */
#define PP_FSTAT_NMIN 3              /* entries [0, NMIN) reduce with min, the rest with max */
#define PP_FSTAT_MAX_EGO_SPEED 3
#define PP_FSTAT_MAX_TARGET_SPEED 4
#define PP_FSTAT_MAX_STEP_SPEED 5    /* max |V_k| */
#define PP_FSTAT_MAX_ACC 6           /* max |A_k|: peak total acceleration along the plans
                                        (acc_T and acc_N of :930-944 combined) */
#define PP_FSTATS_LEN 7
/* fstats_dev[PP_FSTATS_LEN] (device, double, overwritten); asynchronous on cuda_stream. */
int pp_fstats_batch(const pp_plans *plans_dev, int64_t n_frames, double *fstats_dev,
                    void *cuda_stream);

/* ---- the one collective of a multi-GPU job: the final reduction of the statistics ----
 * Frames are independent, so a job shards them contiguously over the GPUs (rank r plans
 * [r N / G, (r + 1) N / G)) and exchanges nothing until every shard is planned; then
 * pp_stats_reduce all-reduces, in place and as one NCCL group on cuda_stream, the int64 vector
 * (ncclSum — exact, so 1/2/4/8-GPU results are identical) and the f64 vector (ncclMin over
 * [0, PP_FSTAT_NMIN), ncclMax over the rest); either pointer may be NULL.  nccl_comm is an
 * ncclComm_t: the caller's own, or one made by the helpers below (one process per GPU: rank 0
 * calls pp_comm_unique_id and ships the PP_COMM_ID_BYTES to the others by any means, every
 * rank calls pp_comm_init_rank with its device current; one process driving several devices:
 * pp_comm_init_all, and the per-device pp_stats_reduce calls between pp_comm_group_begin / _end).
 * NCCL is loaded at run time (the copy already in the process if there is one). */
#define PP_COMM_ID_BYTES 128
int pp_comm_unique_id(void *id_out);
int pp_comm_init_rank(const void *id, int rank, int world, void **nccl_comm_out);
int pp_comm_init_all(int n_dev, const int *devices, void **nccl_comms_out);
int pp_comm_destroy(void *nccl_comm);
int pp_comm_group_begin(void);
int pp_comm_group_end(void);
int pp_stats_reduce(void *nccl_comm, int64_t *stats_dev, double *fstats_dev, void *cuda_stream);

/* pp_plan_batch followed by pp_stats_batch in one call (stats_dev[PP_STATS_LEN], overwritten):
 * the statistics of each chunk are taken as soon as the chunk is planned, concurrently with
 * the planning of the other chunks.  Same results as the two separate calls. */
int pp_plan_stats_batch(const pp_map *map, const pp_config *cfg, const pp_frames *in,
                        const pp_plans *out, int64_t n_frames, int64_t *stats_dev,
                        void *cuda_stream);

/* Kernel selection for pp_plan_batch: 0 = auto (the pipeline; below 1536 frames the
 * warp-per-frame kernel), 1 = one thread per frame in a single kernel, 2 = the pipeline
 * (k_prep, k_cars, k_decide_t, k_emit), 3 = the pipeline with the tiled cars kernel (a warp's
 * frames' cars by TMA into shared memory, reductions in the same kernel), 4 = one warp per
 * frame (lowest latency: closest-waypoint argmin and the per-car work spread over the lanes).
 * All variants produce the same bits. */
int pp_set_kernel_variant(int variant);
/* Chunks of a large batch that pp_plan_batch keeps in flight at once on internal streams:
 * 1..8, 0 = default (4).  1 runs the kernels of the pipeline strictly one after the other,
 * which is what a per-kernel time breakdown needs. */
int pp_set_pipes(int pipes);
/* Measurement aid (bench.py): when on, pp_plan_batch records CUDA events on the
 * caller's stream around each kernel of the pipeline.  pp_get_phase_ms waits for
 * the last of them and returns the summed device time in ms of ms_out[0] = ego
 * preparation, [1] = per-car matching, [2] = decision + spline fit, [3] = point
 * emission, [4] = complete path for the queued rare frames (PP_NUM_PHASES
 * values) over all chunks launched since the previous call (chunks_out, may be
 * NULL). */
#define PP_NUM_PHASES 5
int pp_set_phase_timing(int on);
int pp_get_phase_ms(double *ms_out, int64_t *chunks_out);
/* Number of kernel launches issued by this library since load (bench.py's
 * gpu_launches). */
int64_t pp_launch_count(void);

/* ---- unit-level entry points (device pointers), one per reference function.
 * Each runs the same __device__ code the fused kernel uses. ---- */

/* distancesq_pt_seg (src/helpers.h:188-249): out_d2, out_rnom, out_rdenom, out_snom [n]. */
int pp_distancesq_pt_seg_batch(const double *px, const double *py, const double *ax,
                               const double *ay, const double *bx, const double *by,
                               double *out_d2, double *out_rnom, double *out_rdenom,
                               double *out_snom, int64_t n, void *cuda_stream);

/* Map::init_reference_waypoint (src/main.cpp:143-197) for n points:
 * out_ref_wp[n], out_ratio[n][3]. */
int pp_init_reference_waypoint_batch(const pp_map *map, const double *x, const double *y,
                                     int32_t *out_ref_wp, double *out_ratio, int64_t n,
                                     void *cuda_stream);

/* Map::lane_matching + project_speed (src/main.cpp:199-275,330-358) for n
 * objects, each relative to its own reference point (rx, ry):
 * out_ok, out_lane, out_next_wp [n] ; out_s, out_d, out_vs, out_vd [n]. */
int pp_lane_matching_batch(const pp_map *map, const double *rx, const double *ry,
                           const double *x, const double *y, const double *vx,
                           const double *vy, int32_t *out_ok, int32_t *out_lane,
                           int32_t *out_next_wp, double *out_s, double *out_d,
                           double *out_vs, double *out_vd, int64_t n, void *cuda_stream);

/* Map::project_speed (src/main.cpp:330-358) alone: speed vectors (vx, vy)[n]
 * against the reference-line segment ending at waypoint next_wp[n] (un-wrapped
 * ids as lane_matching returns them): out_vs, out_vd [n]. */
int pp_project_speed_batch(const pp_map *map, const double *vx, const double *vy,
                           const int32_t *next_wp, double *out_vs, double *out_vd, int64_t n,
                           void *cuda_stream);

/* Map::get_lane_pos (src/main.cpp:277-328) relative to (rx, ry):
 * out_x, out_y, out_dist [n]; out_wp [n]. */
int pp_get_lane_pos_batch(const pp_map *map, const double *rx, const double *ry,
                          const double *s, const int32_t *lane, double *out_x,
                          double *out_y, int32_t *out_wp, double *out_dist, int64_t n,
                          void *cuda_stream);

/* tk::spline::set_points + operator() (src/spline.h:284-396): n_splines
 * splines of n_knots (3..15) knots each, knots[k*n_splines... ] laid out
 * spline-major: kx[i*n_knots + k]; each evaluated at n_q query points
 * q[i*n_q + j] -> out[i*n_q + j]. */
int pp_spline_batch(const double *kx, const double *ky, int32_t n_knots, const double *q,
                    int32_t n_q, double *out, int64_t n_splines, void *cuda_stream);

/* Udacity starter helpers (src/helpers.h:43-155); never called by the
 * reference's planner but part of its API surface. maps_* are device arrays
 * of n_wp entries. */
int pp_closest_waypoint_batch(const double *x, const double *y, const double *maps_x,
                              const double *maps_y, int32_t n_wp, int32_t *out, int64_t n,
                              void *cuda_stream);
int pp_next_waypoint_batch(const double *x, const double *y, const double *theta,
                           const double *maps_x, const double *maps_y, int32_t n_wp,
                           int32_t *out, int64_t n, void *cuda_stream);
int pp_get_frenet_batch(const double *x, const double *y, const double *theta,
                        const double *maps_x, const double *maps_y, int32_t n_wp,
                        double *out_s, double *out_d, int64_t n, void *cuda_stream);
int pp_get_xy_batch(const double *s, const double *d, const double *maps_s,
                    const double *maps_x, const double *maps_y, int32_t n_wp, double *out_x,
                    double *out_y, int64_t n, void *cuda_stream);

/* LaneChangePlanner::calculate_target_lane (src/main.cpp:364-485) on explicit,
 * already matched cars: problem i has n_cars cars car_*[i*n_cars + j] (lane < 0
 * = car not in the map) and scalars ego_lane/target_lane/ego_s/ego_vs/dt0 [n].
 * out_target_lane[n]. */
int pp_lane_change_batch(const pp_config *cfg, const int32_t *car_id, const double *car_s,
                         const double *car_vs, const int32_t *car_lane, int32_t n_cars,
                         const int32_t *ego_lane, const int32_t *target_lane, const double *ego_s,
                         const double *ego_vs, const double *dt0, int32_t *out_target_lane,
                         int64_t n, void *cuda_stream);

/* LimitSpeed::calculate (src/main.cpp:1068-1150) for one followed car per
 * element, then SpeedController(ego_speed).add_limit_breakpoint (:495-533):
 * out_ls_* = the LimitSpeed result, out_sc_* = the controller's target after
 * the limit, out_flags = PP_F_COLLISION|BRAKE|MAXBRAKE|ADJUST|KEEP. */
int pp_limit_speed_batch(const pp_config *cfg, const double *car_vx, const double *car_vy,
                         const double *next_s, const double *ego_s, const double *ego_speed,
                         const double *ego_acc, const int32_t *in_lane, double *out_ls_speed,
                         double *out_ls_time, double *out_sc_speed, double *out_sc_time,
                         uint32_t *out_flags, int64_t n, void *cuda_stream);

/* TrajectoryBuilder::build (src/main.cpp:565-1049) on explicit inputs: the ego
 * reference point is the last previous point (or ego_x/ego_y when prev_n < 10),
 * sc_* is the SpeedController state handed to build().  prev_x/prev_y [n][10],
 * out_x/out_y [n][50] (NaN beyond out_n), out_flags [n]. */
int pp_trajectory_build_batch(const pp_map *map, const pp_config *cfg, const int32_t *prev_n,
                              const double *prev_x, const double *prev_y, const double *ego_x,
                              const double *ego_y, const double *ego_yaw_deg,
                              const int32_t *target_lane, const double *ego_d,
                              const double *ego_vd, const double *sc_start,
                              const double *sc_target, const double *sc_time, double *out_x,
                              double *out_y, int32_t *out_n, uint32_t *out_flags, int64_t n,
                              void *cuda_stream);

/* The control points of TrajectoryBuilder::build alone (src/main.cpp:638-768), in the map
 * frame, as the reference logs them (control_points=, :779-781): the start point (pos_x,
 * pos_y) = the last kept previous point, or the telemetry pose on a cold start; then up to 5
 * points on the centre line of target_lane.  sc_start = SpeedController::start_speed (the ego
 * speed).  out_x / out_y [n][6] (quiet NaN beyond out_n[n]). */
int pp_control_points_batch(const pp_map *map, const double *pos_x, const double *pos_y,
                            const int32_t *target_lane, const double *ego_d, const double *ego_vd,
                            const double *sc_start, double *out_x, double *out_y, int32_t *out_n,
                            int64_t n, void *cuda_stream);

/* SpeedController (src/main.cpp:488-548) on n explicit controllers
 * (start, target, time, shift [n], all read; target/time/shift may be rewritten):
 *   op 0  get_speed(a)                 -> out[n]            (:503-512)
 *   op 1  add_limit_breakpoint(a, b)   -> target, time      (:513-533)
 *   op 2  override_speed(a, b)         -> shift             (:534-547)
 *   op 3  SpeedController(start)       -> target, time, shift (:493-501; a, b unused)
 * a, b [n] are the method's arguments; out may be NULL for ops 1-3, a/b for op 3. */
int pp_speed_controller_batch(int32_t op, const double *start, double *target, double *time,
                              double *shift, const double *a, const double *b, double *out,
                              int64_t n, void *cuda_stream);

/* Device-memory helpers, so that a caller (and include/pp.hpp) needs neither
 * the CUDA headers nor libcudart of its own: plain cudaMalloc / cudaFree /
 * cudaMemcpy (synchronous) / cudaDeviceSynchronize on the current device. */
int pp_dev_alloc(void **out, size_t bytes);
int pp_dev_free(void *p);
/* Page-locked host memory (cudaHostAlloc / cudaFreeHost): what the buffers handed to
 * pp_plan_batch_host[_split] should live in — copies from pageable memory are staged by the
 * driver and do not overlap with anything. */
int pp_host_alloc(void **out, size_t bytes);
int pp_host_free(void *p);
int pp_dev_upload(void *dst_dev, const void *src_host, size_t bytes);
int pp_dev_download(void *dst_host, const void *src_dev, size_t bytes);
int pp_dev_sync(void);
/* ... and what a host program driving several devices needs (tools/pp_multi.cpp): make a device
 * current for the calling thread, a stream on the current device (cudaStreamNonBlocking). */
int pp_dev_set(int device);
int pp_stream_create(void **stream_out);
int pp_stream_sync(void *stream);
int pp_stream_destroy(void *stream);

/* Device self-test of the exact-arithmetic helpers the kernels use in place of
 * generic divisions / fmod / atan2 (Markstein quotient with cached reciprocal,
 * x/50, angle wrap, small-slope atan): n random trials; counts_dev[8]:
 * [0..2] = number of results that differ bitwise from the generic operation
 * (must be 0), [3] = largest |fast atan - atan2| seen, in ulps, [4..7] = number
 * of results of the emission kernel's unguarded versions (reciprocal, x/50,
 * atan, angle wrap) that differ from the guarded helpers inside the range their
 * "not covered" flag leaves clear (must be 0). */
int pp_selftest_math(int64_t n, uint64_t seed, int64_t *counts_dev, void *cuda_stream);

/* ---- synthetic workload (host).  Counter-based RNG
 * keyed by (seed, frame index): any sub-range can be generated on any rank.
 * `out` holds HOST pointers (const is cast away; caller-owned). ---- */
int pp_synth_frames(const pp_map *map, uint64_t seed, int64_t first_frame, int64_t n_frames,
                    int32_t n_cars, int32_t rare_permille, const pp_frames *out);
/* The same frames, bit for bit, written by the GPU into DEVICE buffers (asynchronous on
 * cuda_stream): BASELINE config 5 generates its 64M dense frames in HBM, chunk by chunk. */
int pp_synth_frames_dev(const pp_map *map, uint64_t seed, int64_t first_frame, int64_t n_frames,
                        int32_t n_cars, int32_t rare_permille, const pp_frames *out_dev,
                        void *cuda_stream);

/* ---- candidate sweep (BASELINE config 4) --------------------------
 * The reference plans ONE trajectory per frame and notes that "multiple
 * trajectories could be calculated and checked" (README.md:207-208).  The sweep
 * does that: for every frame, PP_SWEEP_CANDS = 3 target lanes x 16 target speeds
 * x 8 target times, each candidate being
 *     SpeedController sc(ego_speed); sc.add_limit_breakpoint(v_i, t_j);
 *     TrajectoryBuilder().build(prev, ..., target_lane = lane, ..., sc)
 * (src/main.cpp:493-533,565-1049) with v_i = i * max_speed / 15, t_j = 0.5 (j + 1) s,
 * candidate index = (lane * 16 + i) * 8 + j.  A candidate is scored on its OUTPUT
 * points P_0..P_{n-1} only (so that the reference's unmodified classes are the
 * oracle), with V_k = (P_{k+1} - P_k) * 50 and A_k = (V_{k+1} - V_k) * 50:
 *     score = |lane - target_lane|                        (target_lane: the planner's own
 *                                                          decision for this frame, :1355-1369)
 *           + (max_speed - mean_k |V_k|) / max_speed
 *           + 0.5 * max(0, max_k |A_k| - maximum_acc)
 * and PP_SWEEP_BAD (1e9) if the builder fell back to the angle-based generator
 * (:848), produced fewer than 3 points, or the score is not a finite number
 * below that (NaN points of a standstill candidate).  The lowest score wins (lowest index on
 * ties) and its trajectory is returned. */
#define PP_SWEEP_LANES 3
#define PP_SWEEP_SPEEDS 16
#define PP_SWEEP_TIMES 8
#define PP_SWEEP_CANDS (PP_SWEEP_LANES * PP_SWEEP_SPEEDS * PP_SWEEP_TIMES)
#define PP_SWEEP_BAD 1e9
typedef struct pp_sweep_out {
  int32_t *best;        /* [N] winning candidate index */
  double *best_score;   /* [N] */
  double *next_x;       /* [N][50] its trajectory (NaN beyond n_points) */
  double *next_y;
  int32_t *n_points;    /* [N] */
  double *scores;       /* [N][PP_SWEEP_CANDS], may be NULL */
} pp_sweep_out;
/* in / out hold DEVICE pointers; asynchronous on cuda_stream. */
int pp_sweep_batch(const pp_map *map, const pp_config *cfg, const pp_frames *in,
                   const pp_sweep_out *out, int64_t n_frames, void *cuda_stream);

/* ---- closed-loop rollouts (BASELINE config 3) -----------------
 * R independent ego vehicles, each driving in closed loop against its own
 * planner output and its own synthetic traffic, entirely on the device:
 *
 *   every tick   frame  <- simulator state        (kernel)
 *                plan   <- pp_plan_batch(frame)    (the pipeline above)
 *                state  <- simulator step(plan)    (kernel)
 *
 * The simulator model stands in for the Udacity term-3 simulator the reference
 * was validated against (README.md:12-17); it is defined here, not in the
 * reference, and restated on the CPU in oracle/ for the parity tests:
 *   - the ego is moved to the consume_k-th point of the returned trajectory, the
 *     remaining points become previous_path (src/main.cpp:1248-1249), target_lane
 *     is carried over (:1195);  speed_mph = distance moved / (0.02 k) * 2.237;
 *     yaw stays at its initial value (the planner only reads it on a cold start);
 *   - each car keeps (lane, segment, ratio, speed) and advances along its lane
 *     centre line at constant speed; x, y, vx, vy are derived from that;
 *   - a car whose ego-centred s (as matched by the planner this tick) leaves
 *     [-100, 300] m, or that the planner dropped, is respawned ahead (200-300 m)
 *     if it fell behind, behind (60-100 m) otherwise, in a random lane at
 *     50 +- 10 mph, from a counter-based generator keyed by (seed, rollout, car, tick).
 * All of it is + - * / only, so the CPU restatement is bit-identical. */
typedef struct pp_rollouts pp_rollouts; /* opaque: owns the device state */

/* Host view of the simulator state (caller-owned arrays, R rollouts, C cars). */
typedef struct pp_rollout_state {
  double *ego_x, *ego_y, *ego_yaw_deg, *ego_speed_mph; /* [R] */
  int32_t *path_n;                                     /* [R] unconsumed points of the last plan */
  double *path_x, *path_y;                             /* [R][50] */
  int32_t *target_lane;                                /* [R] */
  int32_t *car_lane, *car_wp;                          /* [R][C] lane, segment end waypoint (0..n-1) */
  double *car_ratio, *car_speed;                       /* [R][C] position along the segment, m/s */
  int64_t tick;                                        /* ticks simulated so far */
} pp_rollout_state;

/* Initial state r = a pure function of (seed, first_rollout + r): ego at rest on a random lane
 * centre (cold start), n_cars cars within [-100, 300] m.  Built on the host, uploaded. */
int pp_rollouts_create(const pp_map *map, int64_t n_rollouts, int32_t n_cars, uint64_t seed,
                       int64_t first_rollout, pp_rollouts **out);
void pp_rollouts_destroy(pp_rollouts *r);
/* n_ticks closed-loop ticks, consume_k (1..40) points per tick; asynchronous on cuda_stream. */
int pp_rollouts_run(pp_rollouts *r, const pp_config *cfg, int64_t n_ticks, int32_t consume_k,
                    void *cuda_stream);
/* lean != 0: the per-tick plans keep only what the simulator and the statistics read (trajectory,
 * n_points, lanes, ref_wp, flags, per-car lane and s); the diagnostics and the other per-car
 * outputs are not written (their pointers in pp_rollouts_last are NULL).  Default: everything. */
int pp_rollouts_set_lean(pp_rollouts *r, int lean);
/* Stream groups a tick is cut into (1..8, at least 4,096 rollouts each; 0 = automatic: up to 8
 * groups of at least 16,384 rollouts).  The result does not depend on it. */
int pp_rollouts_set_groups(pp_rollouts *r, int groups);
/* Device views of the LAST tick's frames and plans (valid until the next run / destroy). */
int pp_rollouts_last(const pp_rollouts *r, pp_frames *frames_dev, pp_plans *plans_dev);
/* Copy the simulator state out (synchronises the device). */
int pp_rollouts_get_state(const pp_rollouts *r, pp_rollout_state *host_out);
/* Sum over all ticks so far of the per-tick pp_stats_batch vectors: stats_dev[PP_STATS_LEN]
 * (device, int64).  The multi-GPU job all-reduces this. */
int pp_rollouts_stats(const pp_rollouts *r, int64_t *stats_dev, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* PP_B200_H */
