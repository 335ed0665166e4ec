// How should the SMALL per-frame fields of the host entry point cross PCIe?  Copies of the host
// entry point for 1,048,576 frames in 131,072-frame chunks on three round-robin streams
// (profiles/probe_pcie_pattern.py, pattern A), with the 7 small inputs (44 B/frame) and the 5
// small outputs (20 B/frame) moved
//   mode 0: by one cudaMemcpyAsync per field (what pp_plan_batch_host did),
//   mode 1: by one kernel per chunk and direction that reads / writes the pinned host arrays in
//           place (zero copy), the big fields still by the copy engines,
//   mode 2: not at all (the big fields alone: the ceiling of this chunking),
//   mode 3: everything, big fields included, by the in-place kernels (no copy engine at all).
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o probe_pcie_small probe_pcie_small.cu
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::printf("%s: %s\n", #x, cudaGetErrorString(e_));                         \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

struct Ptrs {
  const char *src[7];
  char *dst[7];
  int bytes[7];
  int count;
};

// element-wise copy of `count` arrays of n elements (4 or 8 bytes each), grid-stride
__global__ void k_move(Ptrs p, long n) {
  for (int f = 0; f < p.count; f++) {
    if (p.bytes[f] == 8) {
      const double *s = (const double *)p.src[f];
      double *d = (double *)p.dst[f];
      for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        d[i] = s[i];
    } else {
      const int *s = (const int *)p.src[f];
      int *d = (int *)p.dst[f];
      for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        d[i] = s[i];
    }
  }
}

struct BigPtrs {
  const char *src[9];
  char *dst[9];
  long bytes[9];  // per array, multiples of 16
  int count;
};
__global__ void k_move_big(BigPtrs p) {
  for (int f = 0; f < p.count; f++) {
    const int4 *s = (const int4 *)p.src[f];
    int4 *d = (int4 *)p.dst[f];
    const long n = p.bytes[f] / 16;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
      d[i] = s[i];
  }
}

int main() {
  const long n = 1 << 20, chunk = 131072;
  const int big_in[] = {80, 80, 48, 96, 96, 96, 96}, small_in[] = {8, 8, 8, 8, 4, 4, 4};
  const int big_out[] = {320, 320}, small_out[] = {4, 4, 4, 4, 4};
  std::vector<char *> h_bi(7), h_si(7), h_bo(2), h_so(5);
  char *d_bi[3][7], *d_si[3][7], *d_bo[3][2], *d_so[3][5];
  for (int f = 0; f < 7; f++) CK(cudaHostAlloc((void **)&h_bi[f], n * big_in[f], cudaHostAllocDefault));
  for (int f = 0; f < 7; f++) CK(cudaHostAlloc((void **)&h_si[f], n * small_in[f], cudaHostAllocDefault));
  for (int f = 0; f < 2; f++) CK(cudaHostAlloc((void **)&h_bo[f], n * big_out[f], cudaHostAllocDefault));
  for (int f = 0; f < 5; f++) CK(cudaHostAlloc((void **)&h_so[f], n * small_out[f], cudaHostAllocDefault));
  for (int s = 0; s < 3; s++) {
    for (int f = 0; f < 7; f++) CK(cudaMalloc((void **)&d_bi[s][f], chunk * big_in[f]));
    for (int f = 0; f < 7; f++) CK(cudaMalloc((void **)&d_si[s][f], chunk * small_in[f]));
    for (int f = 0; f < 2; f++) CK(cudaMalloc((void **)&d_bo[s][f], chunk * big_out[f]));
    for (int f = 0; f < 5; f++) CK(cudaMalloc((void **)&d_so[s][f], chunk * small_out[f]));
  }
  cudaStream_t st[3];
  for (int s = 0; s < 3; s++) CK(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
  for (int mode = 0; mode < 4; mode++) {
    for (int blocks : {0, 16, 64, 296, 592}) {
      if ((mode != 1 && mode != 3) != (blocks == 0)) continue;
      if (mode == 1 && blocks == 592) continue;
      double best = 1e9, tot = 0;
      const int reps = 6;
      for (int r = 0; r < reps; r++) {
        CK(cudaDeviceSynchronize());
        const auto t0 = std::chrono::steady_clock::now();
        for (long c = 0; c < n / chunk; c++) {
          const int s = (int)(c % 3);
          const long lo = c * chunk;
          if (mode == 3) {
            BigPtrs b;
            b.count = 7;
            for (int f = 0; f < 7; f++) {
              b.src[f] = h_bi[f] + lo * big_in[f];
              b.dst[f] = d_bi[s][f];
              b.bytes[f] = chunk * big_in[f];
            }
            k_move_big<<<blocks, 256, 0, st[s]>>>(b);
          } else {
            for (int f = 0; f < 7; f++)
              CK(cudaMemcpyAsync(d_bi[s][f], h_bi[f] + lo * big_in[f], chunk * big_in[f],
                                 cudaMemcpyHostToDevice, st[s]));
          }
          if (mode == 0)
            for (int f = 0; f < 7; f++)
              CK(cudaMemcpyAsync(d_si[s][f], h_si[f] + lo * small_in[f], chunk * small_in[f],
                                 cudaMemcpyHostToDevice, st[s]));
          if (mode == 1 || mode == 3) {
            Ptrs p;
            p.count = 7;
            for (int f = 0; f < 7; f++) {
              p.src[f] = h_si[f] + lo * small_in[f];
              p.dst[f] = d_si[s][f];
              p.bytes[f] = small_in[f];
            }
            k_move<<<blocks, 256, 0, st[s]>>>(p, chunk);
          }
          if (mode == 3) {
            BigPtrs b;
            b.count = 2;
            for (int f = 0; f < 2; f++) {
              b.src[f] = d_bo[s][f];
              b.dst[f] = h_bo[f] + lo * big_out[f];
              b.bytes[f] = chunk * big_out[f];
            }
            k_move_big<<<blocks, 256, 0, st[s]>>>(b);
          } else {
            for (int f = 0; f < 2; f++)
              CK(cudaMemcpyAsync(h_bo[f] + lo * big_out[f], d_bo[s][f], chunk * big_out[f],
                                 cudaMemcpyDeviceToHost, st[s]));
          }
          if (mode == 0)
            for (int f = 0; f < 5; f++)
              CK(cudaMemcpyAsync(h_so[f] + lo * small_out[f], d_so[s][f], chunk * small_out[f],
                                 cudaMemcpyDeviceToHost, st[s]));
          if (mode == 1 || mode == 3) {
            Ptrs p;
            p.count = 5;
            for (int f = 0; f < 5; f++) {
              p.src[f] = d_so[s][f];
              p.dst[f] = h_so[f] + lo * small_out[f];
              p.bytes[f] = small_out[f];
            }
            k_move<<<blocks, 256, 0, st[s]>>>(p, chunk);
          }
        }
        CK(cudaDeviceSynchronize());
        const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r > 0) {
          best = t < best ? t : best;
          tot += t;
        }
      }
      std::printf("mode %d blocks %3d: %.2f ms per 1M frames (best %.2f) = %.1f M frames/s\n", mode,
                  blocks, tot / (reps - 1) * 1e3, best * 1e3, n / (tot / (reps - 1)) / 1e6);
    }
  }
  return 0;
}
