// Micro-benchmark: DFMA throughput per SM as a function of resident warps and per-thread ILP,
// plus a mixed DFMA + LDS + integer stream resembling the particle kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_ilp dfma_ilp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: per 4 DFMA one LDS.64 (broadcast) and two integer ops
template <int ILP>
__global__ void kmix(double* out, int iters, double a, double b) {
  __shared__ double sm[256];
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
  double x[ILP];
  int j = threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const double w = sm[(it + u) & 255];
      j = (j * 3 + u) ^ it;
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, w);
    }
  }
  double s = j;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
double run(F f, int blocks, int threads, double flop) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return flop / (ms * 1e-3) / 1e12;
}

int main() {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 2048);
  const int iters = 4096;
  printf("warps/SM  ILP  TFLOP/s (pure DFMA)   TFLOP/s (mixed)\n");
  for (int warps : {4, 8, 16, 24, 32, 64}) {
    const int threads = warps * 32 > 1024 ? 1024 : warps * 32;
    const int blocks = 148 * (warps * 32 / threads);
#define CASE(ILP)                                                                                          \
  {                                                                                                        \
    const double flop = 2.0 * 16 * ILP * (double)iters * blocks * threads;                                \
    double t1 = run([&] { k<ILP><<<blocks, threads>>>(out, iters, 0.999999, 1e-9); }, blocks, threads, flop);   \
    double t2 = run([&] { kmix<ILP><<<blocks, threads>>>(out, iters, 0.999999, 1e-9); }, blocks, threads, flop); \
    printf("%7d  %3d  %8.2f            %8.2f\n", warps, ILP, t1, t2);                                       \
  }
    CASE(1) CASE(2) CASE(4) CASE(8)
  }
  return 0;
}
