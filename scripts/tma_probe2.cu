// Parameterised TMA tensor-load probe: tma_probe2 rank d0 d1 d2 b0 b1 b2 swizzle static_bar
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap pmap, int rank, int elems, int static_bar, int c0, int c1, int c2, float* out) {
  extern __shared__ __align__(1024) unsigned char dyn[];
  __shared__ __align__(8) unsigned long long sbar;
  float* buf = reinterpret_cast<float*>(dyn);
  unsigned long long* barp = static_bar ? &sbar : reinterpret_cast<unsigned long long*>(dyn + ((elems * 4 + 1023) & ~1023));
  const uint32_t b = smem_u32(barp);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(elems * 4) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (rank == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(buf)), "l"(&pmap), "r"(c0), "r"(c1), "r"(b) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(buf)), "l"(&pmap), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
  }
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; ++spin)
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(done) : "r"(b) : "memory");
  float s = 0.f;
  for (int i = threadIdx.x; i < elems; i += blockDim.x) s += buf[i];
  atomicAdd(out, s);
  if (threadIdx.x == 0) out[1] = done ? 1.f : -1.f;
}
int main(int argc, char** argv) {
  const int rank = atoi(argv[1]);
  cuuint64_t dims[3] = {(cuuint64_t)atoi(argv[2]), (cuuint64_t)atoi(argv[3]), (cuuint64_t)atoi(argv[4])};
  cuuint32_t box[3] = {(cuuint32_t)atoi(argv[5]), (cuuint32_t)atoi(argv[6]), (cuuint32_t)atoi(argv[7])};
  const int swz = atoi(argv[8]), static_bar = atoi(argv[9]);
  size_t n = dims[0] * dims[1] * (rank == 3 ? dims[2] : 1);
  std::vector<float> h(n, 1.0f);
  float* vol; cudaMalloc(&vol, n * 4); cudaMemcpy(vol, h.data(), n * 4, cudaMemcpyHostToDevice);
  const cuuint64_t strides[2] = {dims[0] * 4, dims[0] * dims[1] * 4};
  const cuuint32_t ones[3] = {1, 1, 1};
  alignas(64) CUtensorMap map;
  CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, vol, dims, strides, box, ones,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  int elems = box[0] * box[1] * (rank == 3 ? box[2] : 1);
  float* out; cudaMalloc(&out, 8); cudaMemset(out, 0, 8);
  const int smem = ((elems * 4 + 1023) & ~1023) + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int c0 = argc > 10 ? atoi(argv[10]) : 0, c1 = argc > 11 ? atoi(argv[11]) : 0, c2 = argc > 12 ? atoi(argv[12]) : 0;
  probe<<<1, 128, smem>>>(map, rank, elems, static_bar, c0, c1, c2, out);
  cudaError_t e = cudaDeviceSynchronize();
  float res[2] = {0, 0};
  if (e == cudaSuccess) cudaMemcpy(res, out, 8, cudaMemcpyDeviceToHost);
  printf("c0 %d c1 %d c2 %d rank %d dims %llu %llu %llu box %u %u %u swz %d static_bar %d encode=%d -> %s sum=%.0f/%d done=%.0f\n", c0, c1, c2, rank,
         (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1], box[2], swz,
         static_bar, (int)r, cudaGetErrorString(e), res[0], elems, res[1]);
  return 0;
}
