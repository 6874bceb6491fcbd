#include "common.h"

#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mvae_b200.h"

namespace mvae {
namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }
bool pdl_enabled() {   // read at every launch (a getenv, no caching): tests and A/B runs flip it inside one process
  // OPT-IN (MVAE_PDL=1).  Measured on one B200 (profiles/r02_pdl_ab.txt): MNIST B=4096 0.722-0.732 ms/step with the
  // attribute vs 0.711-0.712 without (the early-resident dependents cost more than the hidden launch latency), B=512
  // 0.315-0.316 vs 0.317-0.319 (+0.7 %), FashionMNIST within noise.
  const char* v = getenv("MVAE_PDL");
  return v != nullptr && strcmp(v, "0") != 0;
}
}  // namespace mvae

extern "C" int mvae_version(void) { return 100; }
extern "C" const char* mvae_last_error(void) { return mvae::g_err; }
extern "C" uint64_t mvae_launch_count(void) { return mvae::g_launches.load(std::memory_order_relaxed); }
extern "C" int mvae_device_sm_count(void) {
  static int sms = -1;
  if (sms >= 0) return sms;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  sms = v;
  return sms;
}
