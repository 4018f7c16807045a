// Process-wide bits of the C ABI: error text, version, launch counter, device query.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

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
