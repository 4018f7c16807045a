// Shared helpers for the dgcnn_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dgcnn_b200.h"

namespace dgcnn {

// thread-local last-error text, read through dgcnn_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define DG_REQUIRE(cond, code, ...)                   \
  do {                                                \
    if (!(cond)) return ::dgcnn::set_err((code), __VA_ARGS__); \
  } while (0)

#define DG_CUDA_LAUNCH_CHECK(what)                                                        \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return ::dgcnn::set_err(DGCNN_ERR_CUDA, "%s: %s", (what), cudaGetErrorString(e__)); \
  } while (0)

int num_sms();
void count_launch(int n = 1);
// per-channel fp64 accumulators acc[2][C] filled by the statistics kernels with atomics (edge.cu)
int stats_acc_reset(void* ws, int C, cudaStream_t st);
// pivot (optional, [C]): the totals are those of (z - pivot)
int launch_finalize_stats(const double* acc, int C, double count, float eps, float* mean, float* rstd, cudaStream_t st,
                          const float* pivot = nullptr);
int launch_finalize_sums(const double* acc, int C, float* s1, float* s2, cudaStream_t st);

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace dgcnn
