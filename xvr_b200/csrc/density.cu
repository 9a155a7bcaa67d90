// HU -> density (diffdrr.data.transform_hu_to_density), called on the full volume every training step at
// /root/reference/src/xvr/model/trainer.py:196-197 with a fresh bone multiplier.  The reference makes ~6 passes
// (three masks, masked writes, two reductions, two in-place ops); here the four order statistics the map needs
// are reduced once per HU volume (they do not depend on the multiplier) and the map itself is one read + one write.
#include <string.h>

#include "common.cuh"

namespace xvr {

// monotone float <-> int mapping so that integer atomicMin/Max order floats (deterministic, order independent)
__device__ __forceinline__ int float_key(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float key_float(int k) {
  const int i = k >= 0 ? k : k ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f;
  memcpy(&f, &i, sizeof(f));
  return f;
#endif
}

// keys[0..3] = {min soft, max soft, min bone, max bone} as float keys; caller initialises to {INT_MAX, INT_MIN, ...}
__global__ void __launch_bounds__(256) hu_stats_kernel(const float* __restrict__ hu, int64_t n, float air, float bone,
                                                       int* __restrict__ keys) {
  float smin = INFINITY, smax = -INFINITY, bmin = INFINITY, bmax = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float h = __ldg(hu + i);
    if (h > bone) {
      bmin = fminf(bmin, h);
      bmax = fmaxf(bmax, h);
    } else if (h > air) {
      smin = fminf(smin, h);
      smax = fmaxf(smax, h);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
    smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    bmin = fminf(bmin, __shfl_xor_sync(0xffffffffu, bmin, o));
    bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(keys + 0, float_key(smin));
    atomicMax(keys + 1, float_key(smax));
    atomicMin(keys + 2, float_key(bmin));
    atomicMax(keys + 3, float_key(bmax));
  }
}

__global__ void hu_stats_finish_kernel(const int* __restrict__ keys, float* __restrict__ stats) {
  if (threadIdx.x < 4) stats[threadIdx.x] = key_float(keys[threadIdx.x]);
}

// density = (f(h) - lo) / (hi - lo),  f = soft_min (air), h (soft tissue), m*h (bone)
__global__ void __launch_bounds__(256) hu_map_kernel(const float* __restrict__ hu, int64_t n, float air, float bone,
                                                     float m, const float* __restrict__ m_dev,
                                                     const float* __restrict__ stats, float* __restrict__ out) {
  if (m_dev) m = __ldg(m_dev);  // multiplier from device memory: the launch can sit in a CUDA graph and still vary
  const float smin = stats[0], smax = stats[1], bmin = stats[2], bmax = stats[3];
  const bool has_bone = bmax > -INFINITY;
  float lo = smin, hi = smax;
  if (has_bone) {
    const float a = bmin * m, b = bmax * m;  // m may be negative in principle: take both ends
    lo = fminf(lo, fminf(a, b));
    hi = fmaxf(hi, fmaxf(a, b));
  }
  const float range = hi - lo;  // == max of the shifted map, formed by the same subtraction as its elements
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float h = __ldg(hu + i);
    const float f = h > bone ? h * m : (h > air ? h : smin);
    out[i] = __fdiv_rn(f - lo, range);
  }
}

}  // namespace xvr

using namespace xvr;

// stats (DEVICE, 4 floats) <- {min soft, max soft, min bone, max bone} of hu[n]; workspace: 4 ints (DEVICE)
extern "C" int xvr_hu_stats(const float* hu, long long n, float air, float bone, int* workspace, float* stats,
                            void* stream) {
  if (!hu || !workspace || !stats || n <= 0) {
    set_last_error("xvr_hu_stats: invalid argument");
    return XVR_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int init[4] = {0x7fffffff, (int)0x80000000, 0x7fffffff, (int)0x80000000};
  cudaMemcpyAsync(workspace, init, sizeof(init), cudaMemcpyHostToDevice, st);
  const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  hu_stats_kernel<<<grid, 256, 0, st>>>(hu, n, air, bone, workspace);
  int rc = check_launch("xvr_hu_stats");
  if (rc) return rc;
  hu_stats_finish_kernel<<<1, 32, 0, st>>>(workspace, stats);
  return check_launch("xvr_hu_stats/finish");
}

// multiplier_dev: NULL, or a DEVICE float that overrides `multiplier` (read by the kernel, so a captured launch
// follows the value of the moment)
extern "C" int xvr_hu_to_density(const float* hu, long long n, float air, float bone, float multiplier,
                                 const float* multiplier_dev, const float* stats, float* out, void* stream) {
  if (!hu || !stats || !out || n <= 0) {
    set_last_error("xvr_hu_to_density: invalid argument");
    return XVR_ERR_INVALID;
  }
  const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  hu_map_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hu, n, air, bone, multiplier, multiplier_dev, stats, out);
  return check_launch("xvr_hu_to_density");
}
