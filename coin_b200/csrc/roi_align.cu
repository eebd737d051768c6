// roi_align.cu -- ROIAlign forward / backward and the ROIPooler level rule for sm_100a.
//
// Replaces torchvision::roi_align / torchvision::_roi_align_backward as reached from the reference
// at coin/modeling/roi_heads/clip_roi_heads.py:51-63,142-147,172-176 (ROIPooler -> ROIAlign).
//
// Design (DESIGN.md section "ROIAlign"):
//   * features are read channel-last (fp32 [N,H,W,C]); a warp's 32 lanes are 32 channels (x CPL),
//     so every tap load is one coalesced 128-byte line and the RoI geometry is warp-uniform: all
//     control flow below is branch-uniform and costs no divergence;
//   * one warp owns one output row (k, ph, :) of its channel slab. Along that row the sample
//     x-positions are non-decreasing, so the two feature columns a sample needs are kept in a
//     register window that is advanced (shift + one column load) instead of re-loaded: each
//     feature row is read once per sample row instead of 4 taps x samples per bin;
//   * per bin the samples are still accumulated in the order (iy, ix) with the tap expression
//     ((w1*v1 + w2*v2) + w3*v3) + w4*v4 and un-fused multiply/add (-fmad=false), which is the
//     torchvision CPU kernel's order: forward results are bit-identical to the oracle;
//   * the [channels][bins] slab is transposed through padded shared memory so that the NCHW output
//     ([K,C,PH,PW], 1.2 GB at the benchmark shape - the HBM-bound part) is written as one contiguous
//     coalesced region per CTA;
//   * backward mirrors forward: the register window accumulates tap gradients and is flushed with one
//     fp32 atomic per touched feature cell and channel instead of 4 x samples atomics per bin.
#include "common.cuh"

namespace coin {

struct RoiParams {
    coin_level_t lv[COIN_MAX_LEVELS];
    const float* rois;
    const int32_t* roi_level;
    int C, K, PH, PW, sampling_ratio, aligned;
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
// forward
// ------------------------------------------------------------------------------------------------
template <typename OutT, int PWC, int CPL>
__global__ void __launch_bounds__(512)
roi_align_fwd_kernel(const RoiParams p, OutT* __restrict__ out, const int NBpad, const int chunks) {
    extern __shared__ float tile[];  // [32*CPL][NBpad], NBpad odd -> conflict-free both ways
    const int k = blockIdx.x / chunks;
    const int c0 = (blockIdx.x - k * chunks) * (32 * CPL);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W, C = p.C, PH = p.PH, PW = p.PW;
    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    const float* __restrict__ fbase = L.feat_nhwc + (size_t)g.batch * H * W * C + c0 + lane;
    bool chv[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) chv[j] = (c0 + lane + 32 * j) < C;

    for (int ph = warp; ph < PH; ph += nwarps) {
        for (int pw0 = 0; pw0 < PW; pw0 += PWC) {
            float acc[PWC][CPL];
#pragma unroll
            for (int i = 0; i < PWC; ++i)
#pragma unroll
                for (int j = 0; j < CPL; ++j) acc[i][j] = 0.0f;

            for (int iy = 0; iy < g.grid_h; ++iy) {
                const float y = g.start_h + (float)ph * g.bin_h + ((float)iy + 0.5f) * g.bin_h / (float)g.grid_h;
                int y_lo, y_hi;
                float ly, hy;
                if (!axis_taps(y, H, y_lo, y_hi, ly, hy)) continue;
                const float* __restrict__ rowL = fbase + (size_t)y_lo * W * C;
                const float* __restrict__ rowH = fbase + (size_t)y_hi * W * C;
                int col0 = -1, col1 = -1;  // feature columns currently held in the register window
                float vL0[CPL], vL1[CPL], vH0[CPL], vH1[CPL];
#pragma unroll
                for (int i = 0; i < PWC; ++i) {
                    const int pw = pw0 + i;
                    if (pw >= PW) break;
                    for (int ix = 0; ix < g.grid_w; ++ix) {
                        const float x = g.start_w + (float)pw * g.bin_w + ((float)ix + 0.5f) * g.bin_w / (float)g.grid_w;
                        int x_lo, x_hi;
                        float lx, hx;
                        if (!axis_taps(x, W, x_lo, x_hi, lx, hx)) continue;
                        if (x_lo != col0 || x_hi != col1) {
                            if (x_lo == col1) {
#pragma unroll
                                for (int j = 0; j < CPL; ++j) { vL0[j] = vL1[j]; vH0[j] = vH1[j]; }
                            } else {
#pragma unroll
                                for (int j = 0; j < CPL; ++j) {
                                    vL0[j] = chv[j] ? __ldg(rowL + (size_t)x_lo * C + 32 * j) : 0.0f;
                                    vH0[j] = chv[j] ? __ldg(rowH + (size_t)x_lo * C + 32 * j) : 0.0f;
                                }
                            }
                            if (x_hi == x_lo) {
#pragma unroll
                                for (int j = 0; j < CPL; ++j) { vL1[j] = vL0[j]; vH1[j] = vH0[j]; }
                            } else {
#pragma unroll
                                for (int j = 0; j < CPL; ++j) {
                                    vL1[j] = chv[j] ? __ldg(rowL + (size_t)x_hi * C + 32 * j) : 0.0f;
                                    vH1[j] = chv[j] ? __ldg(rowH + (size_t)x_hi * C + 32 * j) : 0.0f;
                                }
                            }
                            col0 = x_lo;
                            col1 = x_hi;
                        }
                        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
#pragma unroll
                        for (int j = 0; j < CPL; ++j)
                            acc[i][j] += w1 * vL0[j] + w2 * vL1[j] + w3 * vH0[j] + w4 * vH1[j];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < PWC; ++i) {
                if (pw0 + i < PW) {
#pragma unroll
                    for (int j = 0; j < CPL; ++j)
                        tile[(lane + 32 * j) * NBpad + ph * PW + pw0 + i] = acc[i][j] / g.count;
                }
            }
        }
    }
    __syncthreads();

    // The CTA's output region out[k, c0 : c0+cc, :, :] is contiguous: stream it out coalesced.
    const int NB = PH * PW;
    const int cc = min(32 * CPL, C - c0);
    const int total = cc * NB;
    OutT* __restrict__ obase = out + ((size_t)k * C + c0) * NB;
    const int step = blockDim.x;
    const int dc = step / NB, db = step - dc * NB;
    int c = threadIdx.x / NB, b = threadIdx.x - c * NB;
    for (int e = threadIdx.x; e < total; e += step) {
        obase[e] = from_f32<OutT>(tile[c * NBpad + b]);
        c += dc;
        b += db;
        if (b >= NB) { b -= NB; ++c; }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <typename GT, int PWC, int CPL>
__global__ void __launch_bounds__(512)
roi_align_bwd_kernel(const RoiParams p, const GT* __restrict__ grad_out, const int NBpad, const int chunks) {
    extern __shared__ float tile[];
    const int k = blockIdx.x / chunks;
    const int c0 = (blockIdx.x - k * chunks) * (32 * CPL);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W, C = p.C, PH = p.PH, PW = p.PW;
    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    float* __restrict__ gbase = const_cast<float*>(L.feat_nhwc) + (size_t)g.batch * H * W * C + c0 + lane;
    bool chv[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) chv[j] = (c0 + lane + 32 * j) < C;

    {   // stage grad_out[k, c0 : c0+cc, :, :] (contiguous) into the padded tile
        const int NB = PH * PW;
        const int cc = min(32 * CPL, C - c0);
        const int total = cc * NB;
        const GT* __restrict__ ibase = grad_out + ((size_t)k * C + c0) * NB;
        const int step = blockDim.x;
        const int dc = step / NB, db = step - dc * NB;
        int c = threadIdx.x / NB, b = threadIdx.x - c * NB;
        for (int e = threadIdx.x; e < total; e += step) {
            tile[c * NBpad + b] = to_f32(ibase[e]);
            c += dc;
            b += db;
            if (b >= NB) { b -= NB; ++c; }
        }
    }
    __syncthreads();

    for (int ph = warp; ph < PH; ph += nwarps) {
        for (int pw0 = 0; pw0 < PW; pw0 += PWC) {
            float gbin[PWC][CPL];
#pragma unroll
            for (int i = 0; i < PWC; ++i)
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    gbin[i][j] = (pw0 + i < PW && chv[j])
                                     ? tile[(lane + 32 * j) * NBpad + ph * PW + pw0 + i] / g.count
                                     : 0.0f;

            for (int iy = 0; iy < g.grid_h; ++iy) {
                const float y = g.start_h + (float)ph * g.bin_h + ((float)iy + 0.5f) * g.bin_h / (float)g.grid_h;
                int y_lo, y_hi;
                float ly, hy;
                if (!axis_taps(y, H, y_lo, y_hi, ly, hy)) continue;
                float* __restrict__ rowL = gbase + (size_t)y_lo * W * C;
                float* __restrict__ rowH = gbase + (size_t)y_hi * W * C;
                int col0 = -1, col1 = -1;
                float aL0[CPL], aL1[CPL], aH0[CPL], aH1[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) aL0[j] = aL1[j] = aH0[j] = aH1[j] = 0.0f;

                auto flush = [&](const float (&aL)[CPL], const float (&aH)[CPL], int col) {
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        if (chv[j]) {
                            atomicAdd(rowL + (size_t)col * C + 32 * j, aL[j]);
                            atomicAdd(rowH + (size_t)col * C + 32 * j, aH[j]);
                        }
                    }
                };
#pragma unroll
                for (int i = 0; i < PWC; ++i) {
                    const int pw = pw0 + i;
                    if (pw >= PW) break;
                    for (int ix = 0; ix < g.grid_w; ++ix) {
                        const float x = g.start_w + (float)pw * g.bin_w + ((float)ix + 0.5f) * g.bin_w / (float)g.grid_w;
                        int x_lo, x_hi;
                        float lx, hx;
                        if (!axis_taps(x, W, x_lo, x_hi, lx, hx)) continue;
                        if (x_lo != col0 || x_hi != col1) {
                            if (col0 >= 0 && x_lo == col1 && col1 != col0) {
                                flush(aL0, aH0, col0);
#pragma unroll
                                for (int j = 0; j < CPL; ++j) {
                                    aL0[j] = aL1[j]; aH0[j] = aH1[j];
                                    aL1[j] = 0.0f;   aH1[j] = 0.0f;
                                }
                            } else {
                                if (col0 >= 0) {
                                    flush(aL0, aH0, col0);
                                    flush(aL1, aH1, col1);
                                }
#pragma unroll
                                for (int j = 0; j < CPL; ++j) aL0[j] = aL1[j] = aH0[j] = aH1[j] = 0.0f;
                            }
                            col0 = x_lo;
                            col1 = x_hi;
                        }
                        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
#pragma unroll
                        for (int j = 0; j < CPL; ++j) {
                            aL0[j] += gbin[i][j] * w1;
                            aL1[j] += gbin[i][j] * w2;
                            aH0[j] += gbin[i][j] * w3;
                            aH1[j] += gbin[i][j] * w4;
                        }
                    }
                }
                if (col0 >= 0) {
                    flush(aL0, aH0, col0);
                    flush(aL1, aH1, col1);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// layout transforms ([N,C,HW] <-> [N,HW,C]) and the level rule
// ------------------------------------------------------------------------------------------------
template <typename InT>
__global__ void nchw_to_nhwc_kernel(const InT* __restrict__ in, float* __restrict__ out, int C, int HW) {
    __shared__ float t[32][33];
    const size_t img = (size_t)blockIdx.z * C * HW;
    const int hw0 = blockIdx.x * 32, cb = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = cb + r, hw = hw0 + threadIdx.x;
        if (c < C && hw < HW) t[r][threadIdx.x] = to_f32(in[img + (size_t)c * HW + hw]);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int hw = hw0 + r, c = cb + threadIdx.x;
        if (c < C && hw < HW) out[img + (size_t)hw * C + c] = t[threadIdx.x][r];
    }
}

template <typename OutT>
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, OutT* __restrict__ out, int C, int HW) {
    __shared__ float t[32][33];
    const size_t img = (size_t)blockIdx.z * C * HW;
    const int hw0 = blockIdx.x * 32, cb = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int hw = hw0 + r, c = cb + threadIdx.x;
        if (c < C && hw < HW) t[r][threadIdx.x] = in[img + (size_t)hw * C + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = cb + r, hw = hw0 + threadIdx.x;
        if (c < C && hw < HW) out[img + (size_t)c * HW + hw] = from_f32<OutT>(t[threadIdx.x][r]);
    }
}

__global__ void pooler_levels_kernel(const float* __restrict__ boxes, int64_t n, int min_level, int max_level,
                                     float canonical_size, float canonical_level, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = reinterpret_cast<const float4*>(boxes)[i];
    const float size = sqrtf(box_area(b));
    float lvl = floorf(canonical_level + log2f(size / canonical_size + 1e-8f));
    lvl = fminf(fmaxf(lvl, (float)min_level), (float)max_level);
    out[i] = (int32_t)lvl - min_level;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

static int fill_params(RoiParams& p, const coin_level_t* levels, int nlevels, const float* rois,
                       const int32_t* roi_level, int C, int K, int PH, int PW, int sr, int aligned) {
    COIN_REQUIRE(levels && nlevels >= 1 && nlevels <= COIN_MAX_LEVELS, "roi_align: nlevels=%d out of [1,%d]", nlevels, COIN_MAX_LEVELS);
    COIN_REQUIRE(nlevels == 1 || roi_level, "roi_align: roi_level is required when nlevels > 1");
    COIN_REQUIRE(C > 0 && K >= 0 && PH > 0 && PW > 0, "roi_align: bad sizes C=%d K=%d PH=%d PW=%d", C, K, PH, PW);
    COIN_REQUIRE(K == 0 || rois, "roi_align: rois is null");
    for (int i = 0; i < nlevels; ++i) {
        COIN_REQUIRE(levels[i].feat_nhwc && levels[i].H > 0 && levels[i].W > 0, "roi_align: level %d is empty", i);
        p.lv[i] = levels[i];
    }
    for (int i = nlevels; i < COIN_MAX_LEVELS; ++i) p.lv[i] = levels[0];
    p.rois = rois; p.roi_level = nlevels > 1 ? roi_level : nullptr;
    p.C = C; p.K = K; p.PH = PH; p.PW = PW; p.sampling_ratio = sr; p.aligned = aligned;
    return COIN_OK;
}

struct LaunchCfg { int cpl, pwc, threads, nbpad, chunks; size_t smem; };

static int pick_cfg(LaunchCfg& cfg, int C, int PH, int PW, const char* env_prefix) {
    const int NB = PH * PW;
    cfg.nbpad = NB | 1;
    cfg.pwc = (PW > 7) ? 14 : 7;
    int cpl = (cfg.pwc == 14) ? 2 : 4;
    char name[64];
    snprintf(name, sizeof name, "%s_CPL", env_prefix);
    cpl = env_int(name, cpl);
    snprintf(name, sizeof name, "%s_PWC", env_prefix);
    cfg.pwc = env_int(name, cfg.pwc);
    if (cfg.pwc != 7 && cfg.pwc != 14) cfg.pwc = 7;
    if (cpl != 1 && cpl != 2 && cpl != 4) cpl = 2;
    while (cpl > 1 && (32 * (cpl / 2) >= C)) cpl /= 2;  // narrow maps: do not waste lanes-by-channel slots
    while (cpl > 1 && (size_t)32 * cpl * cfg.nbpad * sizeof(float) > 200 * 1024) cpl /= 2;
    cfg.cpl = cpl;
    cfg.smem = (size_t)32 * cpl * cfg.nbpad * sizeof(float);
    if (cfg.smem > 227 * 1024)
        return fail(COIN_ERR_UNSUPPORTED, "roi_align: output %dx%d needs %zu B of shared memory (> 227 KB)", PH, PW, cfg.smem);
    cfg.threads = 32 * (PH < 16 ? PH : 16);
    cfg.chunks = (int)ceil_div(C, 32 * cpl);
    return COIN_OK;
}

template <typename OutT, int PWC, int CPL>
static int launch_fwd(const RoiParams& p, OutT* out, const LaunchCfg& cfg, cudaStream_t s) {
    auto kern = roi_align_fwd_kernel<OutT, PWC, CPL>;
    if (cfg.smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    kern<<<(unsigned)(p.K * cfg.chunks), cfg.threads, cfg.smem, s>>>(p, out, cfg.nbpad, cfg.chunks);
    return check_launch("roi_align_fwd_kernel");
}

template <typename GT, int PWC, int CPL>
static int launch_bwd(const RoiParams& p, const GT* go, const LaunchCfg& cfg, cudaStream_t s) {
    auto kern = roi_align_bwd_kernel<GT, PWC, CPL>;
    if (cfg.smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    kern<<<(unsigned)(p.K * cfg.chunks), cfg.threads, cfg.smem, s>>>(p, go, cfg.nbpad, cfg.chunks);
    return check_launch("roi_align_bwd_kernel");
}

#define COIN_DISPATCH_ROI(FN, T, ptr)                                                          \
    do {                                                                                       \
        if (cfg.pwc == 14) {                                                                   \
            if (cfg.cpl == 4) return FN<T, 14, 4>(p, ptr, cfg, s);                             \
            if (cfg.cpl == 2) return FN<T, 14, 2>(p, ptr, cfg, s);                             \
            return FN<T, 14, 1>(p, ptr, cfg, s);                                               \
        }                                                                                      \
        if (cfg.cpl == 4) return FN<T, 7, 4>(p, ptr, cfg, s);                                  \
        if (cfg.cpl == 2) return FN<T, 7, 2>(p, ptr, cfg, s);                                  \
        return FN<T, 7, 1>(p, ptr, cfg, s);                                                    \
    } while (0)

}  // namespace coin

using namespace coin;

extern "C" int coin_roi_align_fwd(const coin_level_t* levels_host, int nlevels, const float* rois,
                                  const int32_t* roi_level, void* out, int out_dtype, int C, int K, int PH,
                                  int PW, int sampling_ratio, int aligned, coin_stream_t stream) {
    RoiParams p;
    if (int rc = fill_params(p, levels_host, nlevels, rois, roi_level, C, K, PH, PW, sampling_ratio, aligned)) return rc;
    COIN_REQUIRE(out_dtype == COIN_F32 || out_dtype == COIN_F16, "roi_align_fwd: bad out_dtype %d", out_dtype);
    if (K == 0) return COIN_OK;
    COIN_REQUIRE(out, "roi_align_fwd: out is null");
    LaunchCfg cfg;
    if (int rc = pick_cfg(cfg, C, PH, PW, "COIN_ROI_FWD")) return rc;
    cudaStream_t s = as_stream(stream);
    if (out_dtype == COIN_F32) COIN_DISPATCH_ROI(launch_fwd, float, static_cast<float*>(out));
    COIN_DISPATCH_ROI(launch_fwd, __half, static_cast<__half*>(out));
}

extern "C" int coin_roi_align_bwd(const coin_level_t* grad_levels_host, int nlevels, const float* rois,
                                  const int32_t* roi_level, const void* grad_out, int grad_dtype, int C, int K,
                                  int PH, int PW, int sampling_ratio, int aligned, coin_stream_t stream) {
    RoiParams p;
    if (int rc = fill_params(p, grad_levels_host, nlevels, rois, roi_level, C, K, PH, PW, sampling_ratio, aligned)) return rc;
    COIN_REQUIRE(grad_dtype == COIN_F32 || grad_dtype == COIN_F16, "roi_align_bwd: bad grad_dtype %d", grad_dtype);
    if (K == 0) return COIN_OK;
    COIN_REQUIRE(grad_out, "roi_align_bwd: grad_out is null");
    LaunchCfg cfg;
    if (int rc = pick_cfg(cfg, C, PH, PW, "COIN_ROI_BWD")) return rc;
    cudaStream_t s = as_stream(stream);
    if (grad_dtype == COIN_F32) COIN_DISPATCH_ROI(launch_bwd, float, static_cast<const float*>(grad_out));
    COIN_DISPATCH_ROI(launch_bwd, __half, static_cast<const __half*>(grad_out));
}

extern "C" int coin_nchw_to_nhwc_f32(const void* in, int in_dtype, float* out, int N, int C, int H, int W,
                                     coin_stream_t stream) {
    COIN_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad shape");
    COIN_REQUIRE(in_dtype == COIN_F32 || in_dtype == COIN_F16, "nchw_to_nhwc: bad dtype %d", in_dtype);
    if (N == 0) return COIN_OK;
    COIN_REQUIRE(in && out, "nchw_to_nhwc: null pointer");
    const int HW = H * W;
    dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(C, 32), (unsigned)N), block(32, 8);
    if (in_dtype == COIN_F32)
        nchw_to_nhwc_kernel<float><<<grid, block, 0, as_stream(stream)>>>(static_cast<const float*>(in), out, C, HW);
    else
        nchw_to_nhwc_kernel<__half><<<grid, block, 0, as_stream(stream)>>>(static_cast<const __half*>(in), out, C, HW);
    return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int coin_nhwc_f32_to_nchw(const float* in, void* out, int out_dtype, int N, int C, int H, int W,
                                     coin_stream_t stream) {
    COIN_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad shape");
    COIN_REQUIRE(out_dtype == COIN_F32 || out_dtype == COIN_F16, "nhwc_to_nchw: bad dtype %d", out_dtype);
    if (N == 0) return COIN_OK;
    COIN_REQUIRE(in && out, "nhwc_to_nchw: null pointer");
    const int HW = H * W;
    dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(C, 32), (unsigned)N), block(32, 8);
    if (out_dtype == COIN_F32)
        nhwc_to_nchw_kernel<float><<<grid, block, 0, as_stream(stream)>>>(in, static_cast<float*>(out), C, HW);
    else
        nhwc_to_nchw_kernel<__half><<<grid, block, 0, as_stream(stream)>>>(in, static_cast<__half*>(out), C, HW);
    return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int coin_roi_pooler_levels(const float* boxes, int64_t n, int min_level, int max_level,
                                      int canonical_box_size, int canonical_level, int32_t* out_levels,
                                      coin_stream_t stream) {
    COIN_REQUIRE(n >= 0 && min_level <= max_level && canonical_box_size > 0, "roi_pooler_levels: bad arguments");
    if (n == 0) return COIN_OK;
    COIN_REQUIRE(boxes && out_levels, "roi_pooler_levels: null pointer");
    pooler_levels_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        boxes, n, min_level, max_level, (float)canonical_box_size, (float)canonical_level, out_levels);
    return check_launch("pooler_levels_kernel");
}
