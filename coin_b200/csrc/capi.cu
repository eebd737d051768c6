// capi.cu -- error reporting and version of the C ABI (include/coinops.h).
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace coin {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// memset as a KERNEL. cudaMemsetAsync becomes a memset node when a stream is captured into a CUDA graph, and memset
// nodes do not carry the stream's priority: inside the step (coin_b200/pipeline.py) every chain of short kernels that
// contained one queued behind the whole ROIAlign grid of the normal-priority stream. A fill kernel inherits the priority.
__global__ void fill_u32_kernel(uint32_t* __restrict__ p, uint32_t v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

int fill_bytes(void* p, int byte_value, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return COIN_OK;
    if ((reinterpret_cast<uintptr_t>(p) & 3) || (bytes & 3)) return fail(COIN_ERR_INVALID, "fill_bytes: unaligned");
    const uint32_t b = (uint32_t)(byte_value & 0xff);
    const size_t n = bytes / 4;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 4 * 148);
    fill_u32_kernel<<<blocks, 256, 0, s>>>(static_cast<uint32_t*>(p), b * 0x01010101u, n);
    return check_launch("fill_u32_kernel");
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

// Tuning / mode switches (COIN_ROI_EXACT, COIN_ROI_REG, ...). Each name is looked up in the environment ONCE, on first
// use, and cached; coin_set_option overrides it for the process. Launch paths read an int from a small table instead
// of calling getenv() per launch.
namespace {
struct Opt { char name[48]; int value; bool set; };
constexpr int kMaxOpts = 64;
Opt g_opts[kMaxOpts];
int g_nopts = 0;
std::mutex g_opt_mu;
Opt* find_opt(const char* name) {
    for (int i = 0; i < g_nopts; ++i)
        if (strcmp(g_opts[i].name, name) == 0) return &g_opts[i];
    if (g_nopts == kMaxOpts || strlen(name) >= sizeof(g_opts[0].name)) return nullptr;
    Opt* o = &g_opts[g_nopts++];
    strcpy(o->name, name);
    const char* v = getenv(name);
    o->set = v != nullptr;
    o->value = v ? atoi(v) : 0;
    return o;
}
}  // namespace

int option(const char* name, int dflt) {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    Opt* o = find_opt(name);
    return (o && o->set) ? o->value : dflt;
}
}  // namespace coin

extern "C" int coin_set_option(const char* name, int value) {
    if (!name) return coin::fail(COIN_ERR_INVALID, "coin_set_option: null name");
    std::lock_guard<std::mutex> lk(coin::g_opt_mu);
    coin::Opt* o = coin::find_opt(name);
    if (!o) return coin::fail(COIN_ERR_CAPACITY, "coin_set_option: option table full or name too long (%s)", name);
    o->value = value;
    o->set = true;
    return COIN_OK;
}
extern "C" int coin_unset_option(const char* name) {
    if (!name) return coin::fail(COIN_ERR_INVALID, "coin_unset_option: null name");
    std::lock_guard<std::mutex> lk(coin::g_opt_mu);
    coin::Opt* o = coin::find_opt(name);
    if (o) o->set = false;
    return COIN_OK;
}
extern "C" int coin_get_option(const char* name, int dflt) { return name ? coin::option(name, dflt) : dflt; }

extern "C" const char* coin_last_error(void) { return coin::g_err; }
extern "C" int coin_version(void) { return 100; }
extern "C" long long coin_launch_count(void) { return coin::g_launches.load(); }
