// common.cuh -- shared helpers of libcoinops (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/coinops.h"

namespace coin {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// thread-local error message (defined in capi.cu)
int fail(int code, const char* fmt, ...);

void count_launch();  // capi.cu: process-wide counter behind coin_launch_count()
int option(const char* name, int dflt);  // capi.cu: cached environment lookup, overridable by coin_set_option()

inline int check_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(COIN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return COIN_OK;
}

#define COIN_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ::coin::fail(COIN_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// memset as a kernel launch (capi.cu): keeps the launching stream's priority inside captured graphs
int fill_bytes(void* p, int byte_value, size_t bytes, cudaStream_t s);

inline cudaStream_t as_stream(coin_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carve aligned sub-buffers out of a caller-provided workspace.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
    size_t used() const { return align_up(off, 256); }
};

// IoU with detectron2's convention (0 where the intersection is empty). The translation units
// are compiled with -fmad=false, so the multiply/add sequence below is the oracle's, op for op.
__device__ __forceinline__ float box_area(const float4& b) { return (b.z - b.x) * (b.w - b.y); }

__device__ __forceinline__ float iou_d2(const float4& a, float area_a, const float4& b, float area_b) {
    float w = fminf(a.z, b.z) - fmaxf(a.x, b.x);
    float h = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    w = fmaxf(w, 0.0f);
    h = fmaxf(h, 0.0f);
    const float inter = w * h;
    return inter > 0.0f ? inter / (area_a + area_b - inter) : 0.0f;
}

// torchvision nms convention: no guard on an empty intersection (0/x = 0; 0/0 = NaN compares false)
__device__ __forceinline__ float iou_tv(const float4& a, float area_a, const float4& b, float area_b) {
    const float w = fmaxf(0.0f, fminf(a.z, b.z) - fmaxf(a.x, b.x));
    const float h = fmaxf(0.0f, fminf(a.w, b.w) - fmaxf(a.y, b.y));
    const float inter = w * h;
    return inter / (area_a + area_b - inter);
}

}  // namespace coin
