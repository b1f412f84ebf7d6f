#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void atst_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int atst_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    atst_set_error("%s: %s", what, cudaGetErrorString(e));
    return ATST_ERR_CUDA;
  }
  return ATST_OK;
}

extern "C" const char* atst_last_error(void) { return g_err; }
