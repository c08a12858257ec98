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

int num_sms() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

}  // namespace enerf

extern "C" {

const char* enerf_last_error(void) { return enerf::g_err; }
// 2: recomputation (NULL forward_buffer), events / sampler / Adam; 3: CTA cap + scatter CTA size; 4: fp16 shadow in adam_step,
// T_dist in composite_uniform_*; 5: device-side row / ray counts (n_rows_dev, n_alive_dev), grad_mul + fp16 gradients in adam_step,
// torch-topology field, occupancy maintenance
int enerf_abi_version(void) { return 8; }
uint64_t enerf_launch_count(void) { return enerf::g_launches.load(std::memory_order_relaxed); }

}
