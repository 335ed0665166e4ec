// How fast does a B200 SM issue a MIX of FP64 and other instructions?  The planning kernels are
// 32-45 % FP64 (half rate: one warp instruction per ~2.2 cycles per scheduler) and the rest integer
// / move / compare work; their ceiling is neither the FP64 peak nor one instruction per cycle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix profiles/fp64_mix.cu && ./fp64_mix
// Per DFMA, NI independent 32-bit integer multiply-adds (IMAD) and NF FP32 FMAs; ILP chains each.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int NI, int NF>
__global__ void k_mix(double *out, int iters, double a, double b) {
  double v[ILP];
  int w[ILP][NI > 0 ? NI : 1];
  float g[ILP][NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    v[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int k = 0; k < NI; k++) w[i][k] = threadIdx.x + i + k;
#pragma unroll
    for (int k = 0; k < NF; k++) g[i][k] = threadIdx.x * 0.5f + i + k;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      v[i] = fma(v[i], a, b);
#pragma unroll
      for (int k = 0; k < NI; k++) w[i][k] = w[i][k] * 3 + it;
#pragma unroll
      for (int k = 0; k < NF; k++) g[i][k] = fmaf(g[i][k], 1.0000001f, 1e-9f);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    s += v[i];
#pragma unroll
    for (int k = 0; k < NI; k++) s += w[i][k];
#pragma unroll
    for (int k = 0; k < NF; k++) s += g[i][k];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
double run(K kernel, int blocks, int threads, int iters, double *out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kernel<<<blocks, threads>>>(out, 10, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 64 * 1024);
  const int iters = 20000, ILP = 4;
  printf("SMs %d, clock %.0f MHz; per scheduler (SM sub-partition): cycles per group of 1 DFMA + n other\n", sms, khz / 1e3);
  for (int wps : {16, 32}) {
    const int threads = 128, blocks = sms * wps * 32 / threads;
    auto report = [&](const char *name, double secs, int others) {
      // warp instructions per scheduler: warps/SM / 4 schedulers, each iters * ILP groups
      const double groups = (double)wps / 4 * iters * ILP;
      const double cyc = secs * khz * 1e3 / groups;
      printf("  warps/SM %2d  %-22s %.2f cycles per group -> %.2f instr/cycle, FP64 pipe busy %.0f %%\n", wps, name,
             cyc, (1 + others) / cyc, 100 * 2.18 / cyc);
    };
    report("1 DFMA", run(k_mix<ILP, 0, 0>, blocks, threads, iters, out), 0);
    report("1 DFMA + 1 IMAD", run(k_mix<ILP, 1, 0>, blocks, threads, iters, out), 1);
    report("1 DFMA + 2 IMAD", run(k_mix<ILP, 2, 0>, blocks, threads, iters, out), 2);
    report("1 DFMA + 3 IMAD", run(k_mix<ILP, 3, 0>, blocks, threads, iters, out), 3);
    report("1 DFMA + 1 FFMA", run(k_mix<ILP, 0, 1>, blocks, threads, iters, out), 1);
    report("1 DFMA + 2 FFMA", run(k_mix<ILP, 0, 2>, blocks, threads, iters, out), 2);
    report("1 DFMA + 1 IMAD + 1 FFMA", run(k_mix<ILP, 1, 1>, blocks, threads, iters, out), 2);
    report("1 DFMA + 2 IMAD + 2 FFMA", run(k_mix<ILP, 2, 2>, blocks, threads, iters, out), 4);
  }
  return 0;
}
