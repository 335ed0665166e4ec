// FP64 (non-tensor) issue rate of one B200, measured: the compute roofline of this path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak profiles/fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_mix(double *out, int iters, double a, double b) {  // DFMA + integer op pairs
  double v[ILP];
  int w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { v[i] = threadIdx.x * 1e-3 + i; w[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) { v[i] = fma(v[i], a, b); w[i] = w[i] * 3 + it; }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i] + w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
double run(K kernel, int blocks, int threads, int iters, int ilp, double *out) {
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
  return (double)blocks * threads * iters * ilp / (ms * 1e-3);  // thread-level DFMA per second
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 64 * 1024);
  const int iters = 20000;
  printf("SMs %d, clock %.0f MHz\n", sms, khz / 1e3);
  for (int wps : {4, 8, 16, 32, 64}) {  // resident warps per SM
    const int threads = 128, blocks = sms * wps * 32 / threads;
    double r1 = run(k_dfma<1>, blocks, threads, iters, 1, out);
    double r4 = run(k_dfma<4>, blocks, threads, iters, 4, out);
    double r8 = run(k_dfma<8>, blocks, threads, iters, 8, out);
    double m4 = run(k_mix<4>, blocks, threads, iters, 4, out);
    auto per = [&](double r) { return r / sms / (khz * 1e3); };  // DFMA lanes per clock per SM
    printf("warps/SM %2d: DFMA/clk/SM  ilp1 %.1f  ilp4 %.1f  ilp8 %.1f | with 1 int op each: %.1f | peak %.2f TFLOP/s\n",
           wps, per(r1), per(r4), per(r8), per(m4), 2 * r8 / 1e12);
  }
  return 0;
}
