// Shared helpers for the pytristan_b200 C-ABI translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/pytristan_b200.h"

namespace pt {

// thread-local message of the last failure (read through pt_last_error()).
char* last_error_buf();
int fail(int code, const char* fmt, ...);

inline cudaStream_t as_stream(pt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return PT_OK;
}

int sm_count();

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pt

#define PT_REQUIRE(cond, ...) \
  do { if (!(cond)) return pt::fail(PT_ERR_INVALID, __VA_ARGS__); } while (0)
#define PT_CUDA(call) \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return pt::fail(PT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
