// Error reporting and launch accounting for the C ABI (include/bevbert_b200.h).
#include <atomic>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bevbert_b200.h"
#include "common.h"

namespace bb {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(const char* msg) {
  strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return -1;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  char buf[400];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  return set_error(buf);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("BB_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
}  // namespace bb

extern "C" const char* bb_last_error(void) { return bb::g_err; }
extern "C" int bb_abi_version(void) { return 1; }
extern "C" int64_t bb_launch_count(void) { return bb::g_launches.load(); }
extern "C" void bb_reset_launch_count(void) { bb::g_launches.store(0); }

namespace bb {
int set_salt_gemm_tc(const unsigned long long*);
int set_salt_rowops(const unsigned long long*);
int set_salt_attn_flash(const unsigned long long*);
int set_salt_attn_tc(const unsigned long long*);
int set_salt_attn_scores(const unsigned long long*);
int set_salt_gemm_f32(const unsigned long long*);
}  // namespace bb
extern "C" int bb_set_drop_salt_ptr(const uint64_t* device_word) {
  const unsigned long long* p = reinterpret_cast<const unsigned long long*>(device_word);
  int rc = bb::set_salt_gemm_tc(p) | bb::set_salt_rowops(p) | bb::set_salt_attn_flash(p) | bb::set_salt_attn_tc(p) |
           bb::set_salt_attn_scores(p) | bb::set_salt_gemm_f32(p);
  return rc ? bb::set_error("bb_set_drop_salt_ptr: cudaMemcpyToSymbol failed") : 0;
}
