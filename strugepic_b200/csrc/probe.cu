// FP64 FMA throughput probe: the roofline denominator for the W8 particle kernels
// (MEASURED_PEAKS.json has no FP64 entry; SURVEY.md section 8d asks for this probe).
#include <cuda_runtime.h>

#include "../../include/strugepic_b200.h"

namespace {
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

extern "C" int spic_probe_fp64_tflops(int device, double seconds, double* tflops) {
  if (!tflops) return SPIC_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return SPIC_ENODEV;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SPIC_ECUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return SPIC_ECUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);  // warm-up
  cudaDeviceSynchronize();
  const double flop_per_launch = 2.0 * 8 * 16 * (double)iters * blocks * threads;
  double best = 0, elapsed = 0;
  int reps = 0;
  while (elapsed < seconds * 1e3 || reps < 3) {
    cudaEventRecord(e0);
    k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    elapsed += ms;
    ++reps;
    const double tf = flop_per_launch / (ms * 1e-3) / 1e12;
    // sustained figure: report the LAST launch once the requested time has elapsed
    best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return SPIC_OK;
}
