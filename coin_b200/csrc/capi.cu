// capi.cu -- error reporting and version of the C ABI (include/coinops.h).
#include "common.cuh"

#include <atomic>

namespace coin {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace coin

extern "C" const char* coin_last_error(void) { return coin::g_err; }
extern "C" int coin_version(void) { return 100; }
extern "C" long long coin_launch_count(void) { return coin::g_launches.load(); }
