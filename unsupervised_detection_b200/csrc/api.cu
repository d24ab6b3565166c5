// Error plumbing + version of libcis_b200.so (C ABI declared in include/cis_b200.h).
#include "../../include/cis_b200.h"
#include "common.cuh"
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

int cis_set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int cis_set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return CIS_ERR_CUDA;
}
int cis_check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cis_set_cuda_error(e, where);
  return CIS_OK;
}
extern "C" const char* cis_last_error(void) { return g_err; }
extern "C" int cis_version(void) { return 100; }
