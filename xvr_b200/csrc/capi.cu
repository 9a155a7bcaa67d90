// C-ABI plumbing shared by every entry point: error reporting and library identification.
#include <stdio.h>
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

extern "C" int xvr_abi_version(void) { return 2; }

// Number of kernels this library has launched since load (bench.py's gpu_launches evidence).
extern "C" long long xvr_launch_count(void) { return xvr::g_launches; }

// ---------------------------------------------------------------------------------------------- volume texture
// A second, block-linear copy of the CT volume (cudaArray, layered 2D: layer = axis 0 + 1, with one zero layer at
// each end) behind a point-sampled texture object.  The trilinear kernels fetch the 2x2 (axis 1, axis 2) corner footprint of each of the two
// layers with one TLD4 each instead of 8 scalar loads.
extern "C" int xvr_volume_create(int D0, int D1, int D2, void** out) {
  if (!out || D0 < 1 || D1 < 1 || D2 < 1 || D0 > 2046 || D1 > 32768 || D2 > 32768) {
    xvr::set_last_error("xvr_volume_create: invalid shape (layered 2D arrays hold <= 2048 layers of <= 32768^2, "
                        "two of which are the zero padding)");
    return XVR_ERR_INVALID;
  }
  xvr::VolumeTexture* vt = new xvr::VolumeTexture();
  vt->D0 = D0;
  vt->D1 = D1;
  vt->D2 = D2;
  cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
  cudaError_t e = cudaMalloc3DArray(&vt->array, &desc, make_cudaExtent(D2, D1, D0 + 2), cudaArrayLayered);
  if (e == cudaSuccess) {  // zero the two padding layers (0 and D0 + 1) once; uploads never touch them
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
  if (e == cudaSuccess) {
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
  cudaMemcpy3DParms cp = {};
  cp.srcPtr = make_cudaPitchedPtr((void*)volume, (size_t)vt->D2 * sizeof(float), vt->D2, vt->D1);
  cp.dstArray = vt->array;
  cp.dstPos = make_cudaPos(0, 0, 1);  // array layer 0 is zero padding
  cp.extent = make_cudaExtent(vt->D2, vt->D1, vt->D0);
  cp.kind = cudaMemcpyDeviceToDevice;
  cudaError_t e = cudaMemcpy3DAsync(&cp, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    char msg[256];
    snprintf(msg, sizeof(msg), "xvr_volume_upload: %s", cudaGetErrorString(e));
    xvr::set_last_error(msg);
    cudaGetLastError();
    return XVR_ERR_CUDA;
  }
  return XVR_OK;
}

extern "C" int xvr_volume_destroy(void* handle) {
  xvr::VolumeTexture* vt = (xvr::VolumeTexture*)handle;
  if (!vt) return XVR_OK;
  cudaDestroyTextureObject(vt->tex);
  cudaFreeArray(vt->array);
  delete vt;
  return XVR_OK;
}
