// Which tensor-map parameters does UTMALDG accept?  rank, element size, inner box width and tensor size as arguments;
// one case per process (a fault takes the context down).   tma_probe3 rank esize box0 dim0
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, unsigned long long* out, int n8, int rank, int x, int y, int z,
                  unsigned bytes) {
  __shared__ alignas(128) unsigned long long buf[2048];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (rank == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                       smem_u32(buf)), "l"(&map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    else if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                       smem_u32(buf)), "l"(&map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
                       smem_u32(buf)), "l"(&map), "r"(x), "r"(y), "r"(z), "r"(0), "r"(smem_u32(&bar)) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n8; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
  const int rank = atoi(argv[1]), esize = atoi(argv[2]), box0 = atoi(argv[3]), dim0 = atoi(argv[4]);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const int d1 = 10, d2 = 9, d3 = 3;
  const size_t bytes_total = (size_t)dim0 * esize * d1 * d2 * d3;
  std::vector<unsigned char> h(bytes_total);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + (i >> 8));
  unsigned char* d;
  unsigned long long* out;
  cudaMalloc(&d, bytes_total);
  cudaMalloc(&out, 2048 * 8);
  cudaMemcpy(d, h.data(), bytes_total, cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)dim0, d1, d2, d3};
  cuuint64_t strides[3] = {(cuuint64_t)dim0 * esize, (cuuint64_t)dim0 * esize * d1, (cuuint64_t)dim0 * esize * d1 * d2};
  cuuint32_t box[4] = {(cuuint32_t)box0, 4, 4, 3};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_INT32;
  CUresult r = encode(&map, dt, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  size_t boxbytes = (size_t)box0 * esize * 4;
  if (rank >= 3) boxbytes *= 4;
  if (rank >= 4) boxbytes *= 3;
  printf("x0 %d B: rank %d esize %d box0 %d (%d B) dim0 %d: encode rc=%d words %016llx %016llx %016llx %016llx",
         (argc > 5 ? atoi(argv[5]) : (esize == 8 ? 3 : 6)) * esize, rank, esize, box0, box0 * esize, dim0, (int)r, ((unsigned long long*)&map)[0], ((unsigned long long*)&map)[1],
         ((unsigned long long*)&map)[2], ((unsigned long long*)&map)[3]);
  if (r != CUDA_SUCCESS) {
    printf("\n");
    return 1;
  }
  const int x = argc > 5 ? atoi(argv[5]) : (esize == 8 ? 3 : 6);  // default: an odd start in doubles (24 bytes)
  k<<<1, 128>>>(map, out, (int)(boxbytes / 8), rank, x, 2, 1, (unsigned)boxbytes);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  launch: %s\n", cudaGetErrorString(e));
    return 2;
  }
  std::vector<unsigned char> o(boxbytes);
  cudaMemcpy(o.data(), out, boxbytes, cudaMemcpyDeviceToHost);
  size_t bad = 0, t = 0;
  for (int c = 0; c < (rank >= 4 ? 3 : 1); ++c)
    for (int kk = 0; kk < (rank >= 3 ? 4 : 1); ++kk)
      for (int j = 0; j < 4; ++j)
        for (int b = 0; b < box0 * esize; ++b, ++t) {
          const size_t src = (size_t)c * strides[2] + (size_t)(rank >= 3 ? 1 + kk : 0) * strides[1] + (size_t)(2 + j) * strides[0] +
                             (size_t)x * esize + b;
          bad += o[t] != h[src];
        }
  printf("  ok, %zu mismatching bytes\n", bad);
  return 0;
}
