// Process-wide bits of the C ABI: error text, version, launch counter, device query.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"
#include "../../include/dgcnn_b200.h"

namespace dgcnn {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

char* err_buf() { return g_err; }

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int num_sms() {
  static int sms = 0;  // immutable after first query
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else
      return 148;
  }
  return sms;
}

}  // namespace dgcnn

extern "C" int dgcnn_abi_version(void) { return 2; }
extern "C" const char* dgcnn_last_error(void) { return dgcnn::err_buf(); }
extern "C" uint64_t dgcnn_launch_count(void) { return dgcnn::g_launches.load(); }

extern "C" size_t dgcnn_workspace_bytes(int op, int B, int N, int C, int k, int F) {
  (void)k;
  if (B <= 0 || N <= 0) return 0;
  const long long P = (long long)B * N;
  if (P >= (1ll << 31)) return 0;
  switch (op) {
    case DGCNN_OP_KNN: return dgcnn_knn_workspace_bytes(B, N, C);
    case DGCNN_OP_EDGECONV: return dgcnn_edgeconv_workspace_bytes(F);
    case DGCNN_OP_CONV_FWD: return dgcnn_gemm_workspace_bytes((int)P, F, C, 0, 0);
    case DGCNN_OP_CONV_DW: return dgcnn_gemm_workspace_bytes(C, F, (int)P, 1, 0);
    case DGCNN_OP_BN: return dgcnn_bn_workspace_bytes(F);
    case DGCNN_OP_SOFTMAX_XENT: return dgcnn_softmax_xent_workspace_bytes();
    default: return 0;
  }
}
