// roi_common.cuh -- RoI geometry, tap tables and dtype helpers shared by the ROIAlign kernels
// (roi_align.cu: the bit-exact parity kernels and the host entry points; roi_align_sep.cu: the
// separable fast kernels).
#pragma once
#include "common.cuh"

namespace coin {

struct RoiParams {
    coin_level_t lv[COIN_MAX_LEVELS];
    const float* rois;
    const int32_t* roi_level;
    const int32_t* k_dev;   // optional device-side live RoI count (<= K): CTAs of RoIs beyond it exit
    const int32_t* perm;    // optional launch order (coin_roi_launch_order): CTA group i works on RoI perm[i]
    int C, K, PH, PW, sampling_ratio, aligned;
    int flags;              // bit 0: L2-prefetch the next unit's grad_out rows (backward)
};

struct RoiGeom {
    float start_w, start_h, bin_w, bin_h, count;
    int grid_h, grid_w, batch;
};

// Same operation order as oracle/scalar_ref.c::roi_geometry (and the torchvision kernels).
__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi, float scale, int PH,
                                                int PW, int sampling_ratio, int aligned) {
    RoiGeom g;
    g.batch = (int)__ldg(roi);
    const float offset = aligned ? 0.5f : 0.0f;
    g.start_w = __ldg(roi + 1) * scale - offset;
    g.start_h = __ldg(roi + 2) * scale - offset;
    const float end_w = __ldg(roi + 3) * scale - offset;
    const float end_h = __ldg(roi + 4) * scale - offset;
    float roi_w = end_w - g.start_w;
    float roi_h = end_h - g.start_h;
    if (!aligned) {
        roi_w = fmaxf(roi_w, 1.0f);
        roi_h = fmaxf(roi_h, 1.0f);
    }
    g.bin_h = roi_h / (float)PH;
    g.bin_w = roi_w / (float)PW;
    g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / (float)PH);
    g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / (float)PW);
    const int cnt = g.grid_h * g.grid_w;
    g.count = (float)(cnt > 1 ? cnt : 1);
    return g;
}

// One coordinate of a bilinear sample: returns false when the sample lies outside [-1, size].
__device__ __forceinline__ bool axis_taps(float v, int size, int& lo, int& hi, float& l, float& h) {
    if (v < -1.0f || v > (float)size) return false;
    if (v <= 0.0f) v = 0.0f;
    lo = (int)v;
    if (lo >= size - 1) {
        hi = lo = size - 1;
        v = (float)lo;
    } else {
        hi = lo + 1;
    }
    l = v - (float)lo;
    h = 1.0f - l;
    return true;
}

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

// ------------------------------------------------------------------------------------------------
// per-RoI sample tables (shared memory): the x taps depend only on (pw, ix), the y taps only on
// (ph, iy); computing them once per CTA removes ~25 instructions from every sample of every warp.
// ------------------------------------------------------------------------------------------------
struct Tap {          // 16 bytes, read as one LDS.128 (broadcast: every lane reads the same entry)
    int lo, hi;       // element offsets of the low / high cell (x*C or y*W*C); lo < 0: sample outside the map
    float l, h;       // interpolation weights towards hi / lo
};
constexpr int kTapCap = 256;   // entries per axis; larger sampling grids compute taps on the fly

__device__ __forceinline__ Tap make_tap(float start, float bin, int p, int i, int grid, int size, int stride) {
    const float v = start + (float)p * bin + ((float)i + 0.5f) * bin / (float)grid;
    Tap t;
    int lo, hi;
    if (!axis_taps(v, size, lo, hi, t.l, t.h)) {
        t.lo = -1; t.hi = -1; t.l = 0.0f; t.h = 0.0f;
        return t;
    }
    t.lo = lo * stride;
    t.hi = hi * stride;
    return t;
}


// host-side launch of the separable forward kernel (roi_align_sep.cu)
int launch_roi_align_fwd_sep(const RoiParams& p, void* out, int out_dtype, cudaStream_t s);
// host-side launch of the separable backward kernel (roi_align_sep.cu); PW <= 32 only
int launch_roi_align_bwd_sep(const RoiParams& p, const void* grad_out, int grad_dtype, cudaStream_t s);

// register-tile kernels (roi_align_reg.cu): 14x14 / 7x7 outputs, C % 32 == 0, fp32
bool roi_align_fwd_reg_supported(const RoiParams& p, int out_dtype);
int launch_roi_align_fwd_reg(const RoiParams& p, void* out, int out_dtype, cudaStream_t s);
bool roi_align_bwd_reg_supported(const RoiParams& p, int grad_dtype);
int launch_roi_align_bwd_reg(const RoiParams& p, const void* grad_out, int grad_dtype, cudaStream_t s);

}  // namespace coin
