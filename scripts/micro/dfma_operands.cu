// Micro-benchmark: does the FP64 pipe of sm_100 sustain its DFMA rate when the operands are all different
// registers (as in the particle kernels) rather than two loop constants (as in the peak probe)?
// 16 warps per SM (2 blocks x 256 threads, like the kernels), ILP 8 per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8;

// (a) x = fma(x, a, b): a, b loop constants -- what spic_probe_fp64_tflops measures
__global__ void __launch_bounds__(256, 2) k_const(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (b) x[i] = fma(y[i], z[i], x[i]): three different registers per instruction, 24 live doubles
__global__ void __launch_bounds__(256, 2) k_three(double* out, int iters, double a, double b) {
  double x[ILP], y[ILP], z[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = threadIdx.x * 1e-3 + i;
    y[i] = a + 1e-9 * (threadIdx.x + i);
    z[i] = b + 1e-9 * (threadIdx.x - i);
  }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(y[(i + u) % ILP], z[(i + 3 * u + 1) % ILP], x[i]);
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (c) Horner chains: x = fma(x, t, c_k), t a per-thread register, c_k immediates (the weight evaluation)
__global__ void __launch_bounds__(256, 2) k_horner(double* out, int iters, double a, double b) {
  double x[ILP], t[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = threadIdx.x * 1e-3 + i;
    t[i] = a + 1e-9 * (threadIdx.x + i);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      x[i] = fma(x[i], t[i], 0.4375);
      x[i] = fma(x[i], t[i], -0.65625);
      x[i] = fma(x[i], t[i], 0.68359375);
      x[i] = fma(x[i], t[i], 0.1171875);
      x[i] = fma(x[i], t[i], -0.8203125);
      x[i] = fma(x[i], t[i], 0.3828125);
      x[i] = fma(x[i], t[i], 0.658203125);
      x[i] = fma(x[i], t[i], -0.0146484375);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + t[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}

// (d) gather-like: x[i] = fma(w (one LDS.64 broadcast per 4 DFMA), y[i], x[i]) + one IMAD per 4 DFMA
__global__ void __launch_bounds__(256, 2) k_gather(double* out, int iters, double a, double b) {
  __shared__ double sm[256];
  sm[threadIdx.x] = a + 1e-9 * threadIdx.x;
  __syncthreads();
  double x[ILP], y[ILP];
  int j = threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = threadIdx.x * 1e-3 + i;
    y[i] = b + 1e-9 * (threadIdx.x + i);
  }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double w0 = sm[(it + 2 * u) & 255], w1 = sm[(it + 2 * u + 1) & 255];
      j = j * 3 + u;
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(i < 4 ? w0 : w1, y[(i + u) % ILP], x[i]);
    }
  double s = j;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
void run(const char* name, K k, int sms, double* out) {
  const int iters = 4000, blocks = sms * 2, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<<<blocks, threads>>>(out, iters, 0.999999, 1e-3);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<<<blocks, threads>>>(out, iters, 0.999999, 1e-3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * blocks * threads * (double)iters * 8 * ILP;
  std::printf("%-10s %7.3f ms  %6.2f TFLOP/s\n", name, ms, flop / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 2 * 256);
  std::printf("%s, %d SMs, 16 warps per SM, ILP %d\n", p.name, p.multiProcessorCount, ILP);
  run("const", k_const, p.multiProcessorCount, out);
  run("three-reg", k_three, p.multiProcessorCount, out);
  run("horner", k_horner, p.multiProcessorCount, out);
  run("gather", k_gather, p.multiProcessorCount, out);
  return 0;
}
