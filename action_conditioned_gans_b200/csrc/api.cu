// Library-level state of libacg_b200.so: error string, launch counter, device queries.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace acg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("ACG_PDL"); return !e || atoi(e) != 0; }();
    return on;
}

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;  // B200
    }
    return sms;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return ACG_ERR_CUDA;
    }
    count_launch(1);
    return ACG_OK;
}

}  // namespace acg

extern "C" {

int acg_version(void) { return 100; }
const char* acg_last_error(void) { return acg::g_err; }
long long acg_launch_count(void) { return acg::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
