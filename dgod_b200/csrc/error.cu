// Thread-local error string, launch counter and ABI version of libdgod_b200.
#include "common.cuh"

namespace dgod {

static thread_local char t_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

}  // namespace dgod

extern "C" int dgod_abi_version(void) { return DGOD_ABI_VERSION; }
extern "C" const char* dgod_last_error(void) { return dgod::t_error; }
extern "C" uint64_t dgod_launch_count(void) {
  return dgod::g_launches.load(std::memory_order_relaxed);
}
