// C-ABI plumbing shared by every entry point: error reporting and library identification.
#include <stdio.h>
#include <limits.h>
#include <string.h>

#include "common.cuh"

namespace xvr {

static thread_local char g_last_error[512] = "";
static long long g_launches = 0;

void set_last_error(const char* msg) {
  strncpy(g_last_error, msg, sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}

int check_launch(const char* what) {
  __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return XVR_ERR_CUDA;
  }
  return XVR_OK;
}

}  // namespace xvr

extern "C" const char* xvr_last_error(void) { return xvr::g_last_error; }

extern "C" int xvr_abi_version(void) { return 3; }

// Number of kernels this library has launched since load (bench.py's gpu_launches evidence).
extern "C" long long xvr_launch_count(void) { return xvr::g_launches; }

// ---------------------------------------------------------------------------------------------- volume texture
// A second, block-linear copy of the CT volume (cudaArray, layered 2D: layer = axis 0 + 1, with one zero layer at
// each end) behind a point-sampled texture object.  The trilinear kernels fetch the 2x2 (axis 1, axis 2) corner footprint of each of the two
// layers with one TLD4 each instead of 8 scalar loads.
namespace xvr {
// Box of the non-zero voxels, for the marchers' empty-space trimming: CT volumes carry wide margins of air, which
// transform_hu_to_density maps to exactly 0 -- samples whose 8 corners all lie outside this box contribute exact zeros
// to every sum and need not be fetched.  One CTA per (axis-0 index, 8 axis-1 rows); integer atomics: deterministic.
__global__ void volume_bbox_init_kernel(int* bbox, int D0, int D1, int D2) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = threadIdx.x == 0 ? D0 : (threadIdx.x == 1 ? D1 : D2);
  else if (threadIdx.x < 6) bbox[threadIdx.x] = -1;
}
__global__ void __launch_bounds__(256) volume_bbox_kernel(const float* __restrict__ vol, int D1, int D2, int* bbox) {
  __shared__ int s_any, s_ylo, s_yhi, s_zlo, s_zhi;
  if (threadIdx.x == 0) { s_any = 0; s_ylo = INT_MAX; s_yhi = -1; s_zlo = INT_MAX; s_zhi = -1; }
  __syncthreads();
  const int x = blockIdx.y, y = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y < D1) {
    const float* row = vol + ((int64_t)x * D1 + y) * D2;
    int zlo = INT_MAX, zhi = -1;
    for (int z = lane; z < D2; z += 32)
      if (row[z] != 0.f) { zlo = min(zlo, z); zhi = max(zhi, z); }
    zlo = __reduce_min_sync(0xffffffffu, zlo);
    zhi = __reduce_max_sync(0xffffffffu, zhi);
    if (lane == 0 && zhi >= 0) {
      atomicOr(&s_any, 1);
      atomicMin(&s_ylo, y); atomicMax(&s_yhi, y);
      atomicMin(&s_zlo, zlo); atomicMax(&s_zhi, zhi);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_any) {
    atomicMin(bbox + 0, x); atomicMax(bbox + 3, x);
    atomicMin(bbox + 1, s_ylo); atomicMax(bbox + 4, s_yhi);
    atomicMin(bbox + 2, s_zlo); atomicMax(bbox + 5, s_zhi);
  }
}
// ---- Occupancy distance field for the marchers' empty-space trimming (occupied_alpha_range, common.cuh).
// A brick of OCC_BRICK^3 voxels is OCCUPIED if, GROWN BY TWO VOXELS on every side, it holds a non-zero voxel: a sample
// whose cell floor(x) lies in an unoccupied brick reads 8 zero corners (floor(x) + 1 is inside the brick grown by one),
// with one more voxel for rays that graze a brick face within rounding.  What the kernels read is, per brick, the
// CHEBYSHEV DISTANCE (in bricks, capped at OCC_DIST_CAP) to the nearest occupied brick, 0 for an occupied one: a ray in
// a brick at distance D can advance (D - 1) bricks along its fastest axis without meeting an occupied brick.
// Three steps per upload, each voxel read once:
//   1. flags of 2^3-voxel cells (any non-zero voxel): one warp per row of cells, coalesced float2 loads;
//   2. brick occupancy = OR of the cell flags over the brick and one ring of cells around it (= grown by two voxels);
//   3. the distance transform, separable for the max-norm: D(c) = min_dx max(|dx|, min_dy max(|dy|, min_dz:occ |dz|)).
__global__ void __launch_bounds__(256) volume_cell_flags_kernel(const float* __restrict__ vol, int D0, int D1, int D2,
                                                                int m1, int m2, uint8_t* __restrict__ cells) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;  // row of cells: (x2, y2)
  const int x2 = row / m1, y2 = row - x2 * m1;
  if (x2 * 2 >= D0) return;
  for (int z2 = lane; z2 < m2; z2 += 32) {
    int any = 0;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int x = 2 * x2 + dx, y = 2 * y2 + dy;
        if (x < D0 && y < D1) {
          const float* r = vol + ((int64_t)x * D1 + y) * D2 + 2 * z2;
          any |= r[0] != 0.f;
          if (2 * z2 + 1 < D2) any |= r[1] != 0.f;
        }
      }
    cells[((int64_t)x2 * m1 + y2) * m2 + z2] = (uint8_t)any;
  }
}

__global__ void __launch_bounds__(256) volume_brick_occupancy_kernel(const uint8_t* __restrict__ cells, int m0, int m1,
                                                                     int m2, int nb0, int nb1, int nb2,
                                                                     uint8_t* __restrict__ dist) {
  const int b = blockIdx.x * 256 + threadIdx.x;
  if (b >= nb0 * nb1 * nb2) return;
  const int bx = b / (nb1 * nb2), by = (b / nb2) % nb1, bz = b % nb2;
  constexpr int H = OCC_BRICK / 2;  // cells per brick edge
  int any = 0;
  for (int x = max(bx * H - 1, 0); x < min((bx + 1) * H + 1, m0) && !any; ++x)
    for (int y = max(by * H - 1, 0); y < min((by + 1) * H + 1, m1) && !any; ++y)
      for (int z = max(bz * H - 1, 0); z < min((bz + 1) * H + 1, m2); ++z) any |= cells[((int64_t)x * m1 + y) * m2 + z];
  dist[b] = any ? 0 : OCC_DIST_CAP;
}

// one separable pass along `axis` (0, 1, 2): out(c) = min over |o| <= OCC_DIST_CAP of max(|o|, in(c + o e_axis))
__global__ void __launch_bounds__(256) occupancy_distance_pass_kernel(const uint8_t* __restrict__ in,
                                                                      uint8_t* __restrict__ out, int nb0, int nb1,
                                                                      int nb2, int axis) {
  const int b = blockIdx.x * 256 + threadIdx.x;
  if (b >= nb0 * nb1 * nb2) return;
  const int c[3] = {b / (nb1 * nb2), (b / nb2) % nb1, b % nb2};
  const int n = axis == 0 ? nb0 : (axis == 1 ? nb1 : nb2);
  const int stride = axis == 0 ? nb1 * nb2 : (axis == 1 ? nb2 : 1);
  const int i = c[axis];
  int best = in[b];
  for (int o = 1; o < best; ++o) {  // a neighbour |o| away cannot improve on best <= |o|
    if (i - o >= 0) best = min(best, max(o, (int)in[b - o * stride]));
    if (i + o < n) best = min(best, max(o, (int)in[b + o * stride]));
  }
  out[b] = (uint8_t)best;
}
}  // namespace xvr

static int volume_create(int D0, int D1, int D2, void** out, bool with_texture);

extern "C" int xvr_volume_create(int D0, int D1, int D2, void** out) { return volume_create(D0, D1, D2, out, true); }

// The same handle without the texture copy (the Siddon renderer gathers from the linear volume and only wants the
// occupancy the uploads record: no second copy of a 1.8 GB volume).
extern "C" int xvr_occupancy_create(int D0, int D1, int D2, void** out) { return volume_create(D0, D1, D2, out, false); }

static int volume_create(int D0, int D1, int D2, void** out, bool with_texture) {
  if (!out || D0 < 1 || D1 < 1 || D2 < 1 || (with_texture && (D0 > 2046 || D1 > 32768 || D2 > 32768))) {
    xvr::set_last_error("xvr_volume_create: invalid shape (layered 2D arrays hold <= 2048 layers of <= 32768^2, "
                        "two of which are the zero padding)");
    return XVR_ERR_INVALID;
  }
  xvr::VolumeTexture* vt = new xvr::VolumeTexture();
  vt->D0 = D0;
  vt->D1 = D1;
  vt->D2 = D2;
  vt->bbox = nullptr;
  vt->occ = nullptr;
  vt->occ_tmp = nullptr;
  vt->cells = nullptr;
  vt->nb0 = (D0 + xvr::OCC_BRICK - 1) / xvr::OCC_BRICK;
  vt->nb1 = (D1 + xvr::OCC_BRICK - 1) / xvr::OCC_BRICK;
  vt->nb2 = (D2 + xvr::OCC_BRICK - 1) / xvr::OCC_BRICK;
  vt->array = nullptr;
  vt->tex = 0;
  cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
  cudaError_t e = with_texture ? cudaMalloc3DArray(&vt->array, &desc, make_cudaExtent(D2, D1, D0 + 2), cudaArrayLayered)
                               : cudaSuccess;
  if (e == cudaSuccess && with_texture) {  // zero the two padding layers (0 and D0 + 1) once; uploads never touch them
    float* zeros = nullptr;
    e = cudaMalloc(&zeros, (size_t)D1 * D2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(zeros, 0, (size_t)D1 * D2 * sizeof(float));
    for (int layer = 0; layer <= D0 + 1 && e == cudaSuccess; layer += D0 + 1) {
      cudaMemcpy3DParms cp = {};
      cp.srcPtr = make_cudaPitchedPtr(zeros, (size_t)D2 * sizeof(float), D2, D1);
      cp.dstArray = vt->array;
      cp.dstPos = make_cudaPos(0, 0, layer);
      cp.extent = make_cudaExtent(D2, D1, 1);
      cp.kind = cudaMemcpyDeviceToDevice;
      e = cudaMemcpy3D(&cp);
    }
    if (zeros) cudaFree(zeros);
    if (e != cudaSuccess) cudaFreeArray(vt->array);
  }
  if (e == cudaSuccess && with_texture) {
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = vt->array;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    e = cudaCreateTextureObject(&vt->tex, &rd, &td, nullptr);
    if (e != cudaSuccess) cudaFreeArray(vt->array);
  }
  if (e == cudaSuccess) {
    e = cudaMalloc(&vt->bbox, 6 * sizeof(int));
    if (e == cudaSuccess) {  // until the first upload: the whole volume
      const int whole[6] = {0, 0, 0, D0 - 1, D1 - 1, D2 - 1};
      e = cudaMemcpy(vt->bbox, whole, sizeof(whole), cudaMemcpyHostToDevice);
    }
    const size_t nbricks = (size_t)vt->nb0 * vt->nb1 * vt->nb2;
    if (e == cudaSuccess) e = cudaMalloc(&vt->occ, nbricks);
    if (e == cudaSuccess) e = cudaMemset(vt->occ, 0, nbricks);  // until the first upload: every brick occupied
    if (e == cudaSuccess) e = cudaMalloc(&vt->occ_tmp, nbricks);
    if (e == cudaSuccess) e = cudaMalloc(&vt->cells, (size_t)((D0 + 1) / 2) * ((D1 + 1) / 2) * ((D2 + 1) / 2));
    if (e != cudaSuccess) {
      if (vt->tex) cudaDestroyTextureObject(vt->tex);
      if (vt->array) cudaFreeArray(vt->array);
      if (vt->bbox) cudaFree(vt->bbox);
      if (vt->occ) cudaFree(vt->occ);
      if (vt->occ_tmp) cudaFree(vt->occ_tmp);
      if (vt->cells) cudaFree(vt->cells);
    }
  }
  if (e != cudaSuccess) {
    char msg[256];
    snprintf(msg, sizeof(msg), "xvr_volume_create: %s", cudaGetErrorString(e));
    xvr::set_last_error(msg);
    cudaGetLastError();
    delete vt;
    return XVR_ERR_CUDA;
  }
  *out = vt;
  return XVR_OK;
}

extern "C" int xvr_volume_upload(void* handle, const float* volume, void* stream) {
  xvr::VolumeTexture* vt = (xvr::VolumeTexture*)handle;
  if (!vt || !volume) {
    xvr::set_last_error("xvr_volume_upload: null argument");
    return XVR_ERR_INVALID;
  }
  cudaError_t e = cudaSuccess;
  if (vt->array) {
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr((void*)volume, (size_t)vt->D2 * sizeof(float), vt->D2, vt->D1);
    cp.dstArray = vt->array;
    cp.dstPos = make_cudaPos(0, 0, 1);  // array layer 0 is zero padding
    cp.extent = make_cudaExtent(vt->D2, vt->D1, vt->D0);
    cp.kind = cudaMemcpyDeviceToDevice;
    e = cudaMemcpy3DAsync(&cp, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) {
    char msg[256];
    snprintf(msg, sizeof(msg), "xvr_volume_upload: %s", cudaGetErrorString(e));
    xvr::set_last_error(msg);
    cudaGetLastError();
    return XVR_ERR_CUDA;
  }
  xvr::volume_bbox_init_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(vt->bbox, vt->D0, vt->D1, vt->D2);
  xvr::volume_bbox_kernel<<<dim3((vt->D1 + 7) / 8, vt->D0), 256, 0, (cudaStream_t)stream>>>(volume, vt->D1, vt->D2, vt->bbox);
  int rc = xvr::check_launch("xvr_volume_upload/bbox");
  if (rc) return rc;
  {
    cudaStream_t st = (cudaStream_t)stream;
    const int m0 = (vt->D0 + 1) / 2, m1 = (vt->D1 + 1) / 2, m2 = (vt->D2 + 1) / 2;
    const int nbricks = vt->nb0 * vt->nb1 * vt->nb2, grid = (nbricks + 255) / 256;
    xvr::volume_cell_flags_kernel<<<(m0 * m1 + 7) / 8, 256, 0, st>>>(volume, vt->D0, vt->D1, vt->D2, m1, m2, vt->cells);
    xvr::volume_brick_occupancy_kernel<<<grid, 256, 0, st>>>(vt->cells, m0, m1, m2, vt->nb0, vt->nb1, vt->nb2, vt->occ);
    // three passes, ping-pong: occ -> tmp -> occ -> tmp, and the result back into occ (the pointer the kernels hold)
    xvr::occupancy_distance_pass_kernel<<<grid, 256, 0, st>>>(vt->occ, vt->occ_tmp, vt->nb0, vt->nb1, vt->nb2, 2);
    xvr::occupancy_distance_pass_kernel<<<grid, 256, 0, st>>>(vt->occ_tmp, vt->occ, vt->nb0, vt->nb1, vt->nb2, 1);
    xvr::occupancy_distance_pass_kernel<<<grid, 256, 0, st>>>(vt->occ, vt->occ_tmp, vt->nb0, vt->nb1, vt->nb2, 0);
    cudaMemcpyAsync(vt->occ, vt->occ_tmp, (size_t)nbricks, cudaMemcpyDeviceToDevice, st);
  }
  return xvr::check_launch("xvr_volume_upload/occupancy");
}

// The box of non-zero voxels the last upload found: bbox6 HOST int[6] = lo0 lo1 lo2 hi0 hi1 hi2 (synchronises `stream`).
extern "C" int xvr_volume_bbox(void* handle, int* bbox6, void* stream) {
  xvr::VolumeTexture* vt = (xvr::VolumeTexture*)handle;
  if (!vt || !bbox6) {
    xvr::set_last_error("xvr_volume_bbox: null argument");
    return XVR_ERR_INVALID;
  }
  cudaError_t e = cudaMemcpyAsync(bbox6, vt->bbox, 6 * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) {
    xvr::set_last_error(cudaGetErrorString(e));
    cudaGetLastError();
    return XVR_ERR_CUDA;
  }
  return XVR_OK;
}

extern "C" int xvr_volume_destroy(void* handle) {
  xvr::VolumeTexture* vt = (xvr::VolumeTexture*)handle;
  if (!vt) return XVR_OK;
  if (vt->tex) cudaDestroyTextureObject(vt->tex);
  if (vt->array) cudaFreeArray(vt->array);
  if (vt->bbox) cudaFree(vt->bbox);
  if (vt->occ) cudaFree(vt->occ);
  if (vt->occ_tmp) cudaFree(vt->occ_tmp);
  if (vt->cells) cudaFree(vt->cells);
  delete vt;
  return XVR_OK;
}
