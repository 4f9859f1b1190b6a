// Library-wide C-ABI utilities: last-error string and build info.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void eat_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

extern "C" {
const char* eat_last_error(void) { return g_err; }
// ABI version: bump when any signature in include/eat_b200.h changes.
int eat_abi_version(void) { return 3; }
// 0 when a CUDA device of compute capability 10.x is visible, else an EAT_ERR_* code.
int eat_device_check(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) { eat_set_error(cudaGetErrorString(e)); return EAT_ERR_CUDA; }
  if (p.major != 10) { eat_set_error("libeat_b200 is built for sm_100a only"); return EAT_ERR_UNSUPPORTED; }
  return EAT_OK;
}
}
