// pp_comm.cu — what the path does across GPUs, behind the C ABI: the f64 min / max statistics
// of a planned batch and the ONE collective of a multi-GPU job, the final reduction of the
// aggregate statistics (int64 sums with ncclSum, f64 minima / maxima with
// ncclMin / ncclMax; a few hundred bytes, latency bound — NVLink bandwidth plays no role).
//
// NCCL is resolved at run time (dlopen), never at link time: the library loads and plans on a
// box without NCCL, and inside a process that already carries an NCCL (PyTorch bundles its own)
// the same copy is used instead of a second one.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "pp_internal.h"

namespace {

// ---------------------------------------------------------------------------
// f64 statistics: one warp per frame, the 50 points of a row read coalesced.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_f64(double *addr, double v) {
  unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ void atomic_max_f64(double *addr, double v) {
  unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ bool finite_f64(double v) { return fabs(v) <= 1.7976931348623157e308; }

constexpr int kWarps = 8;
__global__ void __launch_bounds__(32 * kWarps)
fstats_kernel(const __grid_constant__ pp_plans p, int64_t n, double *out) {
  __shared__ double s_x[kWarps][PP_PATH_LEN + 2], s_y[kWarps][PP_PATH_LEN + 2];
  __shared__ double s_vx[kWarps][PP_PATH_LEN + 2], s_vy[kWarps][PP_PATH_LEN + 2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double mn[PP_FSTAT_NMIN], mx[PP_FSTATS_LEN - PP_FSTAT_NMIN];
#pragma unroll
  for (int i = 0; i < PP_FSTAT_NMIN; i++) mn[i] = inf;
#pragma unroll
  for (int i = 0; i < PP_FSTATS_LEN - PP_FSTAT_NMIN; i++) mx[i] = -inf;
  const int64_t warps = (int64_t)gridDim.x * kWarps;
  for (int64_t f = (int64_t)blockIdx.x * kWarps + w; f < n; f += warps) {
    const int np = p.n_points[f];
    for (int k = lane; k < PP_PATH_LEN; k += 32) {
      s_x[w][k] = p.next_x[f * PP_PATH_LEN + k];
      s_y[w][k] = p.next_y[f * PP_PATH_LEN + k];
    }
    if (lane == 0) {  // per-frame scalars (optional outputs)
      if (p.ego_speed) {
        const double v = p.ego_speed[f];
        if (finite_f64(v)) {
          mn[PP_FSTAT_MIN_EGO_SPEED] = fmin(mn[PP_FSTAT_MIN_EGO_SPEED], v);
          mx[PP_FSTAT_MAX_EGO_SPEED - PP_FSTAT_NMIN] = fmax(mx[PP_FSTAT_MAX_EGO_SPEED - PP_FSTAT_NMIN], v);
        }
      }
      if (p.target_speed) {
        const double v = p.target_speed[f];
        if (finite_f64(v)) {
          mn[PP_FSTAT_MIN_TARGET_SPEED] = fmin(mn[PP_FSTAT_MIN_TARGET_SPEED], v);
          mx[PP_FSTAT_MAX_TARGET_SPEED - PP_FSTAT_NMIN] =
              fmax(mx[PP_FSTAT_MAX_TARGET_SPEED - PP_FSTAT_NMIN], v);
        }
      }
    }
    __syncwarp();
    // V_k = (P_{k+1} - P_k) * 50, k < np - 1
    for (int k = lane; k < PP_PATH_LEN - 1; k += 32) {
      const double vx = (s_x[w][k + 1] - s_x[w][k]) * 50, vy = (s_y[w][k + 1] - s_y[w][k]) * 50;
      s_vx[w][k] = vx;
      s_vy[w][k] = vy;
      if (k < np - 1) {
        const double sp = sqrt(vx * vx + vy * vy);
        if (finite_f64(sp)) {
          mn[PP_FSTAT_MIN_STEP_SPEED] = fmin(mn[PP_FSTAT_MIN_STEP_SPEED], sp);
          mx[PP_FSTAT_MAX_STEP_SPEED - PP_FSTAT_NMIN] = fmax(mx[PP_FSTAT_MAX_STEP_SPEED - PP_FSTAT_NMIN], sp);
        }
      }
    }
    __syncwarp();
    // A_k = (V_{k+1} - V_k) * 50, k < np - 2
    for (int k = lane; k < PP_PATH_LEN - 2; k += 32) {
      if (k < np - 2) {
        const double ax = (s_vx[w][k + 1] - s_vx[w][k]) * 50, ay = (s_vy[w][k + 1] - s_vy[w][k]) * 50;
        const double a = sqrt(ax * ax + ay * ay);
        if (finite_f64(a)) mx[PP_FSTAT_MAX_ACC - PP_FSTAT_NMIN] = fmax(mx[PP_FSTAT_MAX_ACC - PP_FSTAT_NMIN], a);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < PP_FSTAT_NMIN; i++) {
    double v = mn[i];
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
    if (lane == 0 && v < inf) atomic_min_f64(out + i, v);
  }
#pragma unroll
  for (int i = 0; i < PP_FSTATS_LEN - PP_FSTAT_NMIN; i++) {
    double v = mx[i];
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    if (lane == 0 && v > -inf) atomic_max_f64(out + PP_FSTAT_NMIN + i, v);
  }
}

__global__ void fstats_init(double *out) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  if (threadIdx.x < PP_FSTATS_LEN) out[threadIdx.x] = threadIdx.x < PP_FSTAT_NMIN ? inf : -inf;
}

// ---------------------------------------------------------------------------
// NCCL, resolved at run time.
// ---------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already has
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return;
  g_nccl.handle = h;
#define PP_SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name)
  PP_SYM(GetUniqueId, "ncclGetUniqueId");
  PP_SYM(CommInitRank, "ncclCommInitRank");
  PP_SYM(CommInitAll, "ncclCommInitAll");
  PP_SYM(CommDestroy, "ncclCommDestroy");
  PP_SYM(AllReduce, "ncclAllReduce");
  PP_SYM(GroupStart, "ncclGroupStart");
  PP_SYM(GroupEnd, "ncclGroupEnd");
  PP_SYM(GetErrorString, "ncclGetErrorString");
#undef PP_SYM
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommInitAll && g_nccl.CommDestroy &&
              g_nccl.AllReduce && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.GetErrorString;
}

int need_nccl(const char *who) {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) {
    ppi::set_cuda_error(who, 0, "NCCL (libnccl.so.2) could not be loaded");
    return PP_E_CUDA;
  }
  return PP_OK;
}

int nccl_fail(const char *what, ncclResult_t r) {
  ppi::set_cuda_error(what, (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
  return PP_E_CUDA;
}

}  // namespace

extern "C" int pp_fstats_batch(const pp_plans *p, int64_t n_frames, double *fstats_dev,
                               void *cuda_stream) {
  if (!p || !fstats_dev || n_frames < 0) return PP_E_ARG;
  if (!p->next_x || !p->next_y || !p->n_points) return PP_E_ARG;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  fstats_init<<<1, 32, 0, st>>>(fstats_dev);
  ppi::count_launch();
  if (n_frames > 0) {
    const int64_t want = (n_frames + kWarps - 1) / kWarps;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    fstats_kernel<<<grid, 32 * kWarps, 0, st>>>(*p, n_frames, fstats_dev);
    ppi::count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ppi::set_cuda_error("fstats_kernel", (int)e, cudaGetErrorString(e));
    return PP_E_CUDA;
  }
  return PP_OK;
}

extern "C" int pp_comm_unique_id(void *id_out) {
  if (!id_out) return PP_E_ARG;
  static_assert(sizeof(ncclUniqueId) == PP_COMM_ID_BYTES, "PP_COMM_ID_BYTES");
  int rc = need_nccl("pp_comm_unique_id");
  if (rc != PP_OK) return rc;
  ncclResult_t r = g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId *>(id_out));
  return r == ncclSuccess ? PP_OK : nccl_fail("ncclGetUniqueId", r);
}

extern "C" int pp_comm_init_rank(const void *id, int rank, int world, void **comm_out) {
  if (!id || !comm_out || world < 1 || rank < 0 || rank >= world) return PP_E_ARG;
  *comm_out = nullptr;
  int rc = need_nccl("pp_comm_init_rank");
  if (rc != PP_OK) return rc;
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof uid);
  ncclComm_t c = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&c, world, uid, rank);
  if (r != ncclSuccess) return nccl_fail("ncclCommInitRank", r);
  *comm_out = c;
  return PP_OK;
}

extern "C" int pp_comm_init_all(int n_dev, const int *devices, void **comms_out) {
  if (n_dev < 1 || !comms_out) return PP_E_ARG;
  int rc = need_nccl("pp_comm_init_all");
  if (rc != PP_OK) return rc;
  ncclResult_t r = g_nccl.CommInitAll(reinterpret_cast<ncclComm_t *>(comms_out), n_dev, devices);
  return r == ncclSuccess ? PP_OK : nccl_fail("ncclCommInitAll", r);
}

extern "C" int pp_comm_destroy(void *comm) {
  if (!comm) return PP_OK;
  int rc = need_nccl("pp_comm_destroy");
  if (rc != PP_OK) return rc;
  ncclResult_t r = g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(comm));
  return r == ncclSuccess ? PP_OK : nccl_fail("ncclCommDestroy", r);
}

// The final reduction: int64 sums, f64 minima, f64 maxima, as one NCCL group on the caller's
// stream (in place).  A single-process caller that drives several devices wraps its per-device
// calls in pp_comm_group_begin / pp_comm_group_end.
extern "C" int pp_stats_reduce(void *nccl_comm, int64_t *stats_dev, double *fstats_dev,
                               void *cuda_stream) {
  if (!nccl_comm || (!stats_dev && !fstats_dev)) return PP_E_ARG;
  int rc = need_nccl("pp_stats_reduce");
  if (rc != PP_OK) return rc;
  ncclComm_t c = reinterpret_cast<ncclComm_t>(nccl_comm);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  ncclResult_t r = g_nccl.GroupStart();
  if (r != ncclSuccess) return nccl_fail("ncclGroupStart", r);
  if (stats_dev)
    r = g_nccl.AllReduce(stats_dev, stats_dev, PP_STATS_LEN, ncclInt64, ncclSum, c, st);
  if (r == ncclSuccess && fstats_dev)
    r = g_nccl.AllReduce(fstats_dev, fstats_dev, PP_FSTAT_NMIN, ncclFloat64, ncclMin, c, st);
  if (r == ncclSuccess && fstats_dev)
    r = g_nccl.AllReduce(fstats_dev + PP_FSTAT_NMIN, fstats_dev + PP_FSTAT_NMIN,
                         PP_FSTATS_LEN - PP_FSTAT_NMIN, ncclFloat64, ncclMax, c, st);
  const ncclResult_t r2 = g_nccl.GroupEnd();
  if (r != ncclSuccess) return nccl_fail("ncclAllReduce", r);
  if (r2 != ncclSuccess) return nccl_fail("ncclGroupEnd", r2);
  return PP_OK;
}

extern "C" int pp_comm_group_begin(void) {
  int rc = need_nccl("pp_comm_group_begin");
  if (rc != PP_OK) return rc;
  ncclResult_t r = g_nccl.GroupStart();
  return r == ncclSuccess ? PP_OK : nccl_fail("ncclGroupStart", r);
}
extern "C" int pp_comm_group_end(void) {
  int rc = need_nccl("pp_comm_group_end");
  if (rc != PP_OK) return rc;
  ncclResult_t r = g_nccl.GroupEnd();
  return r == ncclSuccess ? PP_OK : nccl_fail("ncclGroupEnd", r);
}
