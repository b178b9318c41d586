// thread-local error message storage for the C ABI
#include <atomic>
#include "common.cuh"

namespace hsimae {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

// number of kernels this library has launched (bench.py reports it as gpu_launches)
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }


// programmatic dependent launch switch (see common.cuh)
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    v = (getenv("HSIMAE_PDL") && atoi(getenv("HSIMAE_PDL")) == 0) ? 0 : 1;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
void set_pdl(bool on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

}  // namespace hsimae

extern "C" int hsimae_set_pdl(int on) {
  const int was = hsimae::pdl_enabled() ? 1 : 0;
  hsimae::set_pdl(on != 0);
  return was;
}
