// Shared host-side helpers for the octseg C-ABI library (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "octseg.h"

namespace octseg {

char* last_error_buf();                 // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(OCTSEG_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return OCTSEG_OK;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace octseg

#define OCTSEG_CUDA(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return octseg::fail(OCTSEG_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)
