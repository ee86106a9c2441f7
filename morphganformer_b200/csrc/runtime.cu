// runtime.cu -- error text, version, launch counter.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace mgf {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static std::atomic<int> g_fwd_dtype{MGF_BF16};
bool fwd_f16() { return g_fwd_dtype.load(std::memory_order_relaxed) == MGF_F16; }
}  // namespace mgf

extern "C" const char* mgf_last_error(void) { return mgf::g_err; }
extern "C" int mgf_version(void) { return 100; }
extern "C" int64_t mgf_launch_count(void) { return (int64_t)mgf::g_launches.load(); }

extern "C" int mgf_set_forward_dtype(int dtype) {
  if (dtype != MGF_BF16 && dtype != MGF_F16) MGF_FAIL(MGF_E_DTYPE, "set_forward_dtype: only MGF_BF16 or MGF_F16");
  mgf::g_fwd_dtype.store(dtype);
  return 0;
}
extern "C" int mgf_get_forward_dtype(void) { return mgf::g_fwd_dtype.load(); }
