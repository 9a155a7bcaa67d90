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

extern "C" int xvr_abi_version(void) { return 1; }

// Number of kernels this library has launched since load (bench.py's gpu_launches evidence).
extern "C" long long xvr_launch_count(void) { return xvr::g_launches; }
