// common.cuh - shared helpers for libdiffpir_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>

#include "../../include/diffpir_b200.h"

namespace dpir {

// thread-local last error + launch counter (the bench reports how many of OUR kernels ran)
void set_error(const char *fmt, ...);
extern thread_local int64_t g_launch_count;

#define DPIR_CUDA_CHECK(expr)                                                                    \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      dpir::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      (void)cudaGetLastError();                                                                  \
      return DPIR_E_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

#define DPIR_LAUNCH_CHECK()                                                                      \
  do {                                                                                           \
    dpir::g_launch_count++;                                                                      \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess) {                                                                     \
      dpir::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DPIR_E_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

#define DPIR_REQUIRE(cond, ...)                                                                  \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      dpir::set_error(__VA_ARGS__);                                                              \
      return DPIR_E_ARG;                                                                         \
    }                                                                                            \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

}  // namespace dpir
