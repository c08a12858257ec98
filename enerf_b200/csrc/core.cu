// Error state, launch counter and version of the enerf_b200 C-ABI library.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace enerf {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace enerf

extern "C" {

const char* enerf_last_error(void) { return enerf::g_err; }
int enerf_abi_version(void) { return 3; }   // 2: recomputation (NULL forward_buffer), events / sampler / Adam; 3: CTA cap + scatter CTA size
uint64_t enerf_launch_count(void) { return enerf::g_launches.load(std::memory_order_relaxed); }

}
