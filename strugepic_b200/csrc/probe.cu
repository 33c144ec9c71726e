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
// The same 8 chains with THREE distinct register operands per DFMA (x = fma(y, z, x)), which is what the gathers
// and the deposition of the particle kernels issue: the register file feeds such an instruction in 3 cycles per
// scheduler instead of 2 (measured 24.6 against 36.9 TFLOP/s for immediate / reused operands).
__global__ void __launch_bounds__(256) k_dfma3(double* out, int iters, double a, double b) {
  double x[8], y[8], z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = threadIdx.x * 1e-3 + i;
    y[i] = a + 1e-9 * (threadIdx.x + i);
    z[i] = b + 1e-9 * (threadIdx.x - i);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(y[(i + u) % 8], z[(i + 3 * u + 1) % 8], x[i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// Horner chains with IMMEDIATE coefficients, x = fma(x, t, imm): what the weight and line-integral evaluations of
// the particle kernels issue (one register operand pair + a constant): the fastest DFMA form of the three
__global__ void __launch_bounds__(256) k_dfma_imm(double* out, int iters, double t) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, t, 0.25);
      x1 = fma(x1, t, -0.5);
      x2 = fma(x2, t, 0.125);
      x3 = fma(x3, t, -0.75);
      x4 = fma(x4, t, 0.375);
      x5 = fma(x5, t, -0.625);
      x6 = fma(x6, t, 0.875);
      x7 = fma(x7, t, -0.0625);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

static int probe_impl(int device, double seconds, int three_operands, double* tflops);

extern "C" int spic_probe_fp64_tflops(int device, double seconds, double* tflops) {
  return probe_impl(device, seconds, 0, tflops);
}
extern "C" int spic_probe_fp64_three_operand_tflops(int device, double seconds, double* tflops) {
  return probe_impl(device, seconds, 1, tflops);
}
extern "C" int spic_probe_fp64_immediate_tflops(int device, double seconds, double* tflops) {
  return probe_impl(device, seconds, 2, tflops);
}

static int probe_impl(int device, double seconds, int three_operands, double* tflops) {
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
  auto launch = [&]() {
    if (three_operands == 1) k_dfma3<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    else if (three_operands == 2) k_dfma_imm<<<blocks, threads>>>(out, iters, 0.999999);
    else k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
  };
  launch();  // warm-up
  cudaDeviceSynchronize();
  const double flop_per_launch = 2.0 * 8 * 16 * (double)iters * blocks * threads;
  double best = 0, elapsed = 0;
  int reps = 0;
  while (elapsed < seconds * 1e3 || reps < 3) {
    cudaEventRecord(e0);
    launch();
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
