// Minimal probe of the tensor-map forms used by the TMA staging of k_axis_block: which (rank, box, element type)
// combinations does UTMALDG accept for an FP64 guarded field?   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k_probe(const __grid_constant__ CUtensorMap map_param, const CUtensorMap* map_global, double* out, int nout,
                        int c0, int c1, int c2, unsigned bytes) {
  const CUtensorMap* pmap = map_global ? map_global : &map_param;
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 4)
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
              smem_u32(sm)),
          "l"(pmap), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"(smem_u32(&bar))
          : "memory");
    else
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
              smem_u32(sm)),
          "l"(pmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar))
          : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(&bar)), "r"(0)
                 : "memory");
  for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  const int in_global = argc > 2 ? atoi(argv[2]) : 0;  // 1: the descriptor lives in global memory
  const int l2 = argc > 3 ? atoi(argv[3]) : 1;         // 0: no L2 promotion  // one case per process: a fault takes the context down
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no encoder\n");
    return 1;
  }
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const int pj = 12, gy = 10, gz = 9;
  const long pk = pj * gy, pc = pk * gz;
  std::vector<double> h(3 * pc);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&out, 4096 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  struct Case {
    const char* name;
    int rank, dtype, esz;  // dtype 0: FLOAT64, 1: UINT64, 2: UINT32 pairs (inner dim doubled)
    int box0;
  } cases[] = {{"4d f64 box4", 4, 0, 8, 4}, {"4d u64 box4", 4, 1, 8, 4}, {"4d u32 box8", 4, 2, 4, 8},
               {"3d f64 box4", 3, 0, 8, 4}, {"3d u64 box4", 3, 1, 8, 4}, {"4d f64 box2", 4, 0, 8, 2}};
  int idx = -1;
  for (auto& cs : cases) {
    ++idx;
    if (only >= 0 && idx != only) continue;
    CUtensorMap map;
    const int mul = cs.esz == 4 ? 2 : 1;
    cuuint64_t dims[4] = {(cuuint64_t)pj * mul, (cuuint64_t)gy, (cuuint64_t)gz, 3};
    cuuint64_t strides[3] = {(cuuint64_t)pj * 8, (cuuint64_t)pk * 8, (cuuint64_t)pc * 8};
    const int b = cs.box0 / mul;  // box edge in doubles
    cuuint32_t box[4] = {(cuuint32_t)cs.box0, (cuuint32_t)b, (cuuint32_t)b, 3};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMapDataType dt = cs.dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                           : (cs.dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32);
    CUresult r = encode(&map, dt, cs.rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%-14s global=%d l2=%d encode rc=%d", cs.name, in_global, l2, (int)r);
    if (r != CUDA_SUCCESS) {
      printf("\n");
      continue;
    }
    const int ncomp = cs.rank == 4 ? 3 : 1;
    const unsigned bytes = (unsigned)(b * b * b * ncomp * 8);
    const int x0 = 3 * mul, y0 = 2, z0 = 1;  // an ODD x start in doubles: not 16-byte aligned in memory
    CUtensorMap* dmap = nullptr;
    if (in_global) {
      cudaMalloc(&dmap, sizeof map);
      cudaMemcpy(dmap, &map, sizeof map, cudaMemcpyHostToDevice);
    }
    if (cs.rank == 4) k_probe<4><<<1, 64, 4096 * 8>>>(map, dmap, out, b * b * b * ncomp, x0, y0, z0, bytes);
    else k_probe<3><<<1, 64, 4096 * 8>>>(map, dmap, out, b * b * b, x0, y0, z0, bytes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("  launch: %s\n", cudaGetErrorString(e));
      return 2;  // (the context is gone)
    }
    std::vector<double> o(b * b * b * ncomp);
    cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < ncomp; ++c)
      for (int k = 0; k < b; ++k)
        for (int j = 0; j < b; ++j)
          for (int i = 0; i < b; ++i) {
            const double want = (double)(c * pc + (z0 + k) * pk + (y0 + j) * pj + 3 + i);
            if (o[((c * b + k) * b + j) * b + i] != want) ++bad;
          }
    printf("  ok, %d mismatches\n", bad);
  }
  return 0;
}
