// Probe: can sm_100a gather the 2x2 (y,z) footprint of a layered fp32 texture with tld4.a2d, and how fast is a
// ray-march that fetches its 8 trilinear corners that way compared with 8 scalar LDGs?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe_tex scripts/probe_tex.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e)); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ float4 gather_a2d(cudaTextureObject_t tex, int layer, float u, float v) {
  float4 r;
  asm volatile("tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%7}];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(tex), "r"(layer), "f"(u), "f"(v));
  return r;
}

__global__ void check_kernel(cudaTextureObject_t tex, int D0, int D1, int D2, const float* vol, int* bad, float* first) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  // pseudo-random voxel, including borders
  unsigned h = t * 2654435761u;
  int ix = h % D0;
  int iy = (int)((h >> 8) % (D1 + 1)) - 1;
  int iz = (int)((h >> 17) % (D2 + 1)) - 1;
  float4 g = gather_a2d(tex, ix, (float)iz + 1.0f, (float)iy + 1.0f);
  auto at = [&](int y, int z) -> float {
    if (y < 0 || y >= D1 || z < 0 || z >= D2) return 0.f;
    return vol[((size_t)ix * D1 + y) * D2 + z];
  };
  // expected order: x=(u0,v1) y=(u1,v1) z=(u1,v0) w=(u0,v0), u=z axis, v=y axis
  float e0 = at(iy + 1, iz), e1 = at(iy + 1, iz + 1), e2 = at(iy, iz + 1), e3 = at(iy, iz);
  if (t == 0) { first[0] = g.x; first[1] = g.y; first[2] = g.z; first[3] = g.w; first[4] = e0; first[5] = e1; first[6] = e2; first[7] = e3; }
  if (g.x != e0 || g.y != e1 || g.z != e2 || g.w != e3) atomicAdd(bad, 1);
}

struct Geo {
  float s[3];
  float o[3], u[3], v[3];  // target(i,j) = o + i*u + j*v (voxel coords)
  int H, W, np, D0, D1, D2;
};

__device__ __forceinline__ bool ray_range(const Geo& g, int i, int j, float d[3], float& amin, float& amax) {
  amin = 0.f; amax = 1.f;
  const int dims[3] = {g.D0 - 1, g.D1 - 1, g.D2 - 1};
  for (int a = 0; a < 3; ++a) {
    float t = g.o[a] + i * g.u[a] + j * g.v[a];
    d[a] = t - g.s[a] + 1e-8f;
    float a0 = (0.f - g.s[a]) / d[a], a1 = ((float)dims[a] - g.s[a]) / d[a];
    amin = fmaxf(amin, fminf(a0, a1));
    amax = fminf(amax, fmaxf(a0, a1));
  }
  return amin < amax;
}

template <int MODE>  // 0: 8 LDG, 1: 2 tld4, 2: tex3D-like single fetch per sample is not possible on a2d -> 1 tld4 only (rate probe)
__global__ void __launch_bounds__(256) march_kernel(Geo g, const float* __restrict__ vol, cudaTextureObject_t tex, float* out, int lw) {
  int tiles_x = g.W >> 4;
  int tile = blockIdx.x;
  int ty = tile / tiles_x, tx = tile % tiles_x;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int lj = lane & ((1 << lw) - 1), li = lane >> lw;
  int wpr = 16 >> lw;
  int wj = warp % wpr, wi = warp / wpr;
  int j = tx * 16 + wj * (1 << lw) + lj;
  int i = ty * 16 + wi * (32 >> lw) + li;
  float d[3], amin, amax;
  float acc = 0.f;
  if (ray_range(g, i, j, d, amin, amax)) {
    float span = amax - amin, st = 1.f / (g.np - 1);
    const int s0 = g.D1 * g.D2, s1 = g.D2;
#pragma unroll 4
    for (int k = 0; k < g.np; ++k) {
      float al = amin + span * (k * st);
      float x = g.s[0] + al * d[0], y = g.s[1] + al * d[1], z = g.s[2] + al * d[2];
      float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
      int ix = min(max((int)fx0, 0), g.D0 - 2), iy = min(max((int)fy0, 0), g.D1 - 2), iz = min(max((int)fz0, 0), g.D2 - 2);
      float fx = x - fx0, fy = y - fy0, fz = z - fz0;
      float c000, c001, c010, c011, c100, c101, c110, c111;
      if (MODE == 0) {
        const float* p = vol + ((size_t)ix * s0 + iy * s1 + iz);
        c000 = __ldg(p); c001 = __ldg(p + 1); c010 = __ldg(p + s1); c011 = __ldg(p + s1 + 1);
        c100 = __ldg(p + s0); c101 = __ldg(p + s0 + 1); c110 = __ldg(p + s0 + s1); c111 = __ldg(p + s0 + s1 + 1);
      } else {
        float4 a = gather_a2d(tex, ix, (float)iz + 1.0f, (float)iy + 1.0f);
        float4 b = MODE == 1 ? gather_a2d(tex, ix + 1, (float)iz + 1.0f, (float)iy + 1.0f) : a;
        c010 = a.x; c011 = a.y; c001 = a.z; c000 = a.w;
        c110 = b.x; c111 = b.y; c101 = b.z; c100 = b.w;
      }
      float c00 = c000 + fz * (c001 - c000), c01 = c010 + fz * (c011 - c010);
      float c10 = c100 + fz * (c101 - c100), c11 = c110 + fz * (c111 - c110);
      float c0 = c00 + fy * (c01 - c00), c1 = c10 + fy * (c11 - c10);
      acc += c0 + fx * (c1 - c0);
    }
  }
  out[(size_t)blockIdx.y * g.H * g.W + i * g.W + j] = acc;
}

int main() {
  const int N = 512, D0 = N, D1 = N, D2 = N;
  size_t nvox = (size_t)D0 * D1 * D2;
  std::vector<float> h(nvox);
  unsigned s = 12345;
  for (size_t i = 0; i < nvox; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) * (1.0f / 16777216.f); }
  float* vol;
  CK(cudaMalloc(&vol, nvox * 4));
  CK(cudaMemcpy(vol, h.data(), nvox * 4, cudaMemcpyHostToDevice));

  cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
  cudaArray_t arr = nullptr;
  cudaExtent ext = make_cudaExtent(D2, D1, D0);
  cudaError_t e = cudaMalloc3DArray(&arr, &desc, ext, cudaArrayLayered | cudaArrayTextureGather);
  printf("cudaMalloc3DArray(Layered|TextureGather): %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) {
    cudaGetLastError();
    e = cudaMalloc3DArray(&arr, &desc, ext, cudaArrayLayered);
    printf("cudaMalloc3DArray(Layered): %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  cudaMemcpy3DParms cp = {};
  cp.srcPtr = make_cudaPitchedPtr(vol, D2 * 4, D2, D1);
  cp.dstArray = arr;
  cp.extent = ext;
  cp.kind = cudaMemcpyDeviceToDevice;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaMemcpy3D(&cp));
  CK(cudaEventRecord(e0));
  CK(cudaMemcpy3D(&cp));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("linear -> layered array copy of %zu MiB: %.3f ms\n", nvox * 4 >> 20, ms);

  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t tex;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));

  int* bad;
  float* first;
  CK(cudaMalloc(&bad, 4));
  CK(cudaMalloc(&first, 32));
  CK(cudaMemset(bad, 0, 4));
  check_kernel<<<4096, 256>>>(tex, D0, D1, D2, vol, bad, first);
  CK(cudaDeviceSynchronize());
  int hbad;
  float hf[8];
  CK(cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hf, first, 32, cudaMemcpyDeviceToHost));
  printf("gather check: %d mismatches of %d; first got (%g %g %g %g) expected (%g %g %g %g)\n", hbad, 4096 * 256, hf[0], hf[1], hf[2], hf[3], hf[4], hf[5], hf[6], hf[7]);

  // geometry: source 1600 voxels from the centre along -y, detector 440 voxels behind the centre, 256^2 pixels of
  // 2.18 voxels, with a pose rotation about z and x
  const int B = 116, H = 256, W = 256;
  float* out;
  CK(cudaMalloc(&out, (size_t)B * H * W * 4));
  for (int cfg = 0; cfg < 3; ++cfg) {
    float az = cfg == 0 ? 0.f : cfg == 1 ? 0.5f : 0.3f, ax = cfg == 0 ? 0.f : cfg == 1 ? 0.0f : 0.6f, roll = cfg == 2 ? 0.2f : 0.f;
    Geo g;
    g.H = H; g.W = W; g.np = 500; g.D0 = D0; g.D1 = D1; g.D2 = D2;
    float c = (N - 1) / 2.f;
    // camera basis: view dir w, right r (detector j), up q (detector i)
    float w[3] = {sinf(az) * cosf(ax), cosf(az) * cosf(ax), sinf(ax)};
    float r0[3] = {cosf(az), -sinf(az), 0.f};
    float q0[3] = {w[1] * r0[2] - w[2] * r0[1], w[2] * r0[0] - w[0] * r0[2], w[0] * r0[1] - w[1] * r0[0]};
    float r[3], q[3];
    for (int a = 0; a < 3; ++a) { r[a] = cosf(roll) * r0[a] + sinf(roll) * q0[a]; q[a] = -sinf(roll) * r0[a] + cosf(roll) * q0[a]; }
    float pix = 2.1764f;
    for (int a = 0; a < 3; ++a) {
      g.s[a] = c - 1600.f * w[a];
      g.u[a] = pix * q[a];
      g.v[a] = pix * r[a];
      g.o[a] = c + 440.f * w[a] - 127.5f * g.u[a] - 127.5f * g.v[a];
    }
    for (int lw = 0; lw <= 4; lw += (cfg == 0 ? 1 : 2)) {
      for (int mode = 0; mode < 3; ++mode) {
        dim3 grid((H / 16) * (W / 16), B);
        auto launch = [&]() {
          if (mode == 0) march_kernel<0><<<grid, 256>>>(g, vol, tex, out, lw);
          if (mode == 1) march_kernel<1><<<grid, 256>>>(g, vol, tex, out, lw);
          if (mode == 2) march_kernel<2><<<grid, 256>>>(g, vol, tex, out, lw);
        };
        launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        launch();
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        float hsum;
        CK(cudaMemcpy(&hsum, out + 128 * 256 + 128, 4, cudaMemcpyDeviceToHost));
        printf("cfg %d lanes %dx%d mode %s: %.3f ms per 116 DRRs (centre pixel %.4f)\n", cfg, 32 >> lw, 1 << lw,
               mode == 0 ? "8xLDG " : mode == 1 ? "2xTLD4" : "1xTLD4", ms / 2, hsum);
      }
    }
  }
  return 0;
}
