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
//   * the x taps (per pw, ix) and y taps (per ph, iy) are computed once per CTA into shared-memory
//     tables with the oracle's exact operation order, so sample positions are bit-identical;
//   * per bin the samples are accumulated in the order (iy, ix). COIN_ROI_EXACT=1 (parity mode) keeps
//     the tap expression ((w1*v1 + w2*v2) + w3*v3) + w4*v4 un-fused (-fmad=false) and is bit-identical
//     to the torchvision CPU kernel; the default accumulates the same terms with FMAs (half the
//     arithmetic instructions, ~1e-7 relative difference, inside the 1e-5 parity tolerance);
//   * the [channels][bins] slab is assembled in shared memory in the exact layout of the CTA's
//     contiguous output region out[k, c0:c0+CC, :, :] and leaves through 1-D TMA bulk stores
//     (cp.async.bulk.global.shared::cta) - no per-element store instructions for the 1.2 GB output;
//   * backward mirrors forward: the register window accumulates tap gradients and is flushed with one
//     fp32 atomic per touched feature cell and channel instead of 4 x samples atomics per bin.
#include "roi_common.cuh"

namespace coin {

// ---- TMA bulk copies (shared <-> global, 1-D) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store_g2s_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// EXACT: tap expression and accumulation order of the torchvision CPU kernel with un-fused
// multiply/add -> bit-identical to the oracle. !EXACT: the same samples accumulated with FMAs
// (4 instead of 8 instructions per channel and sample, ~1e-7 relative difference).
template <typename OutT, int PWC, int CPL, bool EXACT>
__global__ void __launch_bounds__(256, (CPL >= 4 ? 2 : 3))
roi_align_fwd_kernel(const RoiParams p, OutT* __restrict__ out, const int chunks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ Tap xs[kTapCap], ys[kTapCap];
    OutT* tile = reinterpret_cast<OutT*>(smem_raw);  // [32*CPL][PH*PW] == the CTA's contiguous output region
    const int k = blockIdx.x / chunks;
    if (p.k_dev && k >= __ldg(p.k_dev)) return;   // capacity launch: RoI beyond the live count
    const int c0 = (blockIdx.x - k * chunks) * (32 * CPL);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W, C = p.C, PH = p.PH, PW = p.PW, NB = PH * PW;
    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    const float* __restrict__ fbase = L.feat_nhwc + (size_t)g.batch * H * W * C + c0 + lane;
    bool chv[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) chv[j] = (c0 + lane + 32 * j) < C;

    const bool tabx = PW * g.grid_w <= kTapCap, taby = PH * g.grid_h <= kTapCap;
    if (tabx)
        for (int s = threadIdx.x; s < PW * g.grid_w; s += blockDim.x) {
            const int pw = s / g.grid_w;
            xs[s] = make_tap(g.start_w, g.bin_w, pw, s - pw * g.grid_w, g.grid_w, W, C);
        }
    if (taby)
        for (int s = threadIdx.x; s < PH * g.grid_h; s += blockDim.x) {
            const int ph = s / g.grid_h;
            ys[s] = make_tap(g.start_h, g.bin_h, ph, s - ph * g.grid_h, g.grid_h, H, W * C);
        }
    __syncthreads();

    // 1/count: exact for powers of two (and skipped for 1); the EXACT path divides otherwise
    const int icount = (int)g.count;
    const bool pow2 = (icount & (icount - 1)) == 0;
    const float rcount = 1.0f / g.count;

    for (int ph = warp; ph < PH; ph += nwarps) {
        for (int pw0 = 0; pw0 < PW; pw0 += PWC) {
            float acc[PWC][CPL];
#pragma unroll
            for (int i = 0; i < PWC; ++i)
#pragma unroll
                for (int j = 0; j < CPL; ++j) acc[i][j] = 0.0f;

            for (int iy = 0; iy < g.grid_h; ++iy) {
                const Tap Y = taby ? ys[ph * g.grid_h + iy] : make_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H, W * C);
                if (Y.lo < 0) continue;
                const float* __restrict__ rowL = fbase + Y.lo;
                const float* __restrict__ rowH = fbase + Y.hi;
                const float ly = Y.l, hy = Y.h;
                int col0 = -1, col1 = -1;  // element offsets of the columns held in the register window
                float vL0[CPL], vL1[CPL], vH0[CPL], vH1[CPL];
#pragma unroll
                for (int i = 0; i < PWC; ++i) {
                    const int pw = pw0 + i;
                    if (pw >= PW) break;
                    for (int ix = 0; ix < g.grid_w; ++ix) {
                        const Tap X = tabx ? xs[pw * g.grid_w + ix] : make_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W, C);
                        if (X.lo < 0) continue;
                        if (X.lo != col0 || X.hi != col1) {
                            if (X.lo == col1) {
#pragma unroll
                                for (int j = 0; j < CPL; ++j) { vL0[j] = vL1[j]; vH0[j] = vH1[j]; }
                            } else {
#pragma unroll
                                for (int j = 0; j < CPL; ++j) {
                                    vL0[j] = chv[j] ? __ldg(rowL + X.lo + 32 * j) : 0.0f;
                                    vH0[j] = chv[j] ? __ldg(rowH + X.lo + 32 * j) : 0.0f;
                                }
                            }
                            if (X.hi == X.lo) {
#pragma unroll
                                for (int j = 0; j < CPL; ++j) { vL1[j] = vL0[j]; vH1[j] = vH0[j]; }
                            } else {
#pragma unroll
                                for (int j = 0; j < CPL; ++j) {
                                    vL1[j] = chv[j] ? __ldg(rowL + X.hi + 32 * j) : 0.0f;
                                    vH1[j] = chv[j] ? __ldg(rowH + X.hi + 32 * j) : 0.0f;
                                }
                            }
                            col0 = X.lo;
                            col1 = X.hi;
                        }
                        const float w1 = hy * X.h, w2 = hy * X.l, w3 = ly * X.h, w4 = ly * X.l;
#pragma unroll
                        for (int j = 0; j < CPL; ++j) {
                            if (EXACT) {
                                acc[i][j] += w1 * vL0[j] + w2 * vL1[j] + w3 * vH0[j] + w4 * vH1[j];
                            } else {
                                acc[i][j] = __fmaf_rn(w1, vL0[j], acc[i][j]);
                                acc[i][j] = __fmaf_rn(w2, vL1[j], acc[i][j]);
                                acc[i][j] = __fmaf_rn(w3, vH0[j], acc[i][j]);
                                acc[i][j] = __fmaf_rn(w4, vH1[j], acc[i][j]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < PWC; ++i) {
                if (pw0 + i < PW) {
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        float v = acc[i][j];
                        if (icount != 1) v = (pow2 || !EXACT) ? v * rcount : v / g.count;
                        tile[(lane + 32 * j) * NB + ph * PW + pw0 + i] = from_f32<OutT>(v);
                    }
                }
            }
        }
    }

    // The CTA's output region out[k, c0 : c0+cc, :, :] is contiguous: hand it to the TMA as 1-D bulk
    // stores (no per-element store instructions); fall back to a coalesced copy when unaligned.
    const int cc = min(32 * CPL, C - c0);
    const uint32_t bytes = (uint32_t)cc * NB * sizeof(OutT);
    OutT* __restrict__ obase = out + ((size_t)k * C + c0) * NB;
    if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(obase) & 15u) == 0) {
        bulk_store_g2s_fence();
        __syncthreads();
        if (threadIdx.x < 4) {
            const uint32_t piece = ((bytes / 4 + 15u) / 16u) * 16u;
            const uint32_t off = piece * threadIdx.x;
            if (off < bytes) {
                bulk_store(reinterpret_cast<char*>(obase) + off, smem_raw + off, min(piece, bytes - off));
                bulk_store_commit_wait();
            }
        }
    } else {
        __syncthreads();
        for (int e = threadIdx.x; e < cc * NB; e += blockDim.x) obase[e] = tile[e];
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <typename GT, int PWC, int CPL>
__global__ void __launch_bounds__(256, (CPL >= 4 ? 2 : 3))
roi_align_bwd_kernel(const RoiParams p, const GT* __restrict__ grad_out, const int chunks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ Tap xs[kTapCap], ys[kTapCap];
    __shared__ __align__(8) uint64_t mbar;
    const GT* tile = reinterpret_cast<const GT*>(smem_raw);  // [32*CPL][PH*PW]
    const int k = blockIdx.x / chunks;
    if (p.k_dev && k >= __ldg(p.k_dev)) return;   // capacity launch: RoI beyond the live count
    const int c0 = (blockIdx.x - k * chunks) * (32 * CPL);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W, C = p.C, PH = p.PH, PW = p.PW, NB = PH * PW;
    const int cc = min(32 * CPL, C - c0);
    const uint32_t bytes = (uint32_t)cc * NB * sizeof(GT);
    const GT* __restrict__ ibase = grad_out + ((size_t)k * C + c0) * NB;
    const bool bulk = (bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(ibase) & 15u) == 0;

    // stage grad_out[k, c0 : c0+cc, :, :] (contiguous) into shared memory, asynchronously
    if (bulk) {
        if (threadIdx.x == 0) mbar_init(&mbar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&mbar, bytes);
            const uint32_t piece = ((bytes / 4 + 15u) / 16u) * 16u;
            for (uint32_t off = 0; off < bytes; off += piece)
                bulk_load(smem_raw + off, reinterpret_cast<const char*>(ibase) + off, min(piece, bytes - off), &mbar);
        }
    } else {
        GT* wtile = reinterpret_cast<GT*>(smem_raw);
        for (int e = threadIdx.x; e < cc * NB; e += blockDim.x) wtile[e] = ibase[e];
    }

    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    float* __restrict__ gbase = const_cast<float*>(L.feat_nhwc) + (size_t)g.batch * H * W * C + c0 + lane;
    bool chv[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) chv[j] = (c0 + lane + 32 * j) < C;
    const bool tabx = PW * g.grid_w <= kTapCap, taby = PH * g.grid_h <= kTapCap;
    if (tabx)
        for (int s = threadIdx.x; s < PW * g.grid_w; s += blockDim.x) {
            const int pw = s / g.grid_w;
            xs[s] = make_tap(g.start_w, g.bin_w, pw, s - pw * g.grid_w, g.grid_w, W, C);
        }
    if (taby)
        for (int s = threadIdx.x; s < PH * g.grid_h; s += blockDim.x) {
            const int ph = s / g.grid_h;
            ys[s] = make_tap(g.start_h, g.bin_h, ph, s - ph * g.grid_h, g.grid_h, H, W * C);
        }
    __syncthreads();
    if (bulk) mbar_wait(&mbar, 0);
    const float rcount = 1.0f / g.count;

    for (int ph = warp; ph < PH; ph += nwarps) {
        for (int pw0 = 0; pw0 < PW; pw0 += PWC) {
            float gbin[PWC][CPL];
#pragma unroll
            for (int i = 0; i < PWC; ++i)
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    gbin[i][j] = (pw0 + i < PW && chv[j]) ? to_f32(tile[(lane + 32 * j) * NB + ph * PW + pw0 + i]) * rcount : 0.0f;

            for (int iy = 0; iy < g.grid_h; ++iy) {
                const Tap Y = taby ? ys[ph * g.grid_h + iy] : make_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H, W * C);
                if (Y.lo < 0) continue;
                float* __restrict__ rowL = gbase + Y.lo;
                float* __restrict__ rowH = gbase + Y.hi;
                const float ly = Y.l, hy = Y.h;
                int col0 = -1, col1 = -1;
                float aL0[CPL], aL1[CPL], aH0[CPL], aH1[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) aL0[j] = aL1[j] = aH0[j] = aH1[j] = 0.0f;

                auto flush = [&](const float (&aL)[CPL], const float (&aH)[CPL], int col) {
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        if (chv[j]) {
                            atomicAdd(rowL + col + 32 * j, aL[j]);
                            atomicAdd(rowH + col + 32 * j, aH[j]);
                        }
                    }
                };
#pragma unroll
                for (int i = 0; i < PWC; ++i) {
                    const int pw = pw0 + i;
                    if (pw >= PW) break;
                    for (int ix = 0; ix < g.grid_w; ++ix) {
                        const Tap X = tabx ? xs[pw * g.grid_w + ix] : make_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W, C);
                        if (X.lo < 0) continue;
                        if (X.lo != col0 || X.hi != col1) {
                            if (col0 >= 0 && X.lo == col1 && col1 != col0) {
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
                            col0 = X.lo;
                            col1 = X.hi;
                        }
                        const float w1 = hy * X.h, w2 = hy * X.l, w3 = ly * X.h, w4 = ly * X.l;
#pragma unroll
                        for (int j = 0; j < CPL; ++j) {
                            aL0[j] = __fmaf_rn(gbin[i][j], w1, aL0[j]);
                            aL1[j] = __fmaf_rn(gbin[i][j], w2, aL1[j]);
                            aH0[j] = __fmaf_rn(gbin[i][j], w3, aH0[j]);
                            aH1[j] = __fmaf_rn(gbin[i][j], w4, aH1[j]);
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

// The same transforms with 128 (hw) x 32 (c) tiles and every load of a thread issued before the barrier: 16 independent
// loads per thread instead of 4, a quarter of the CTAs (the 32 x 32 version leaves the copy at ~3.2 TB/s on a 68 MB round trip).
constexpr int kLtHw = 128;
template <typename InT>
__global__ void __launch_bounds__(256) nchw_to_nhwc_v2_kernel(const InT* __restrict__ in, float* __restrict__ out, int C, int HW) {
    __shared__ float t[32][kLtHw + 1];
    const size_t img = (size_t)blockIdx.z * C * HW;
    const int hw0 = blockIdx.x * kLtHw, cb = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;      // (32, 8)
    float v[4][kLtHw / 32];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < kLtHw / 32; ++j) {
            const int c = cb + ty + 8 * i, hw = hw0 + tx + 32 * j;
            v[i][j] = (c < C && hw < HW) ? to_f32(in[img + (size_t)c * HW + hw]) : 0.0f;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < kLtHw / 32; ++j) t[ty + 8 * i][tx + 32 * j] = v[i][j];
    __syncthreads();
    const int c = cb + tx;
#pragma unroll 4
    for (int r = ty; r < kLtHw; r += 8) {
        const int hw = hw0 + r;
        if (c < C && hw < HW) out[img + (size_t)hw * C + c] = t[tx][r];
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256) nhwc_to_nchw_v2_kernel(const float* __restrict__ in, OutT* __restrict__ out, int C, int HW) {
    __shared__ float t[kLtHw][33];
    const size_t img = (size_t)blockIdx.z * C * HW;
    const int hw0 = blockIdx.x * kLtHw, cb = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = cb + tx;
    float v[kLtHw / 8];
#pragma unroll
    for (int i = 0; i < kLtHw / 8; ++i) {
        const int hw = hw0 + ty + 8 * i;
        v[i] = (c < C && hw < HW) ? in[img + (size_t)hw * C + c] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kLtHw / 8; ++i) t[ty + 8 * i][tx] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < kLtHw / 32; ++j) {
            const int cc = cb + ty + 8 * i, hw = hw0 + tx + 32 * j;
            if (cc < C && hw < HW) out[img + (size_t)cc * HW + hw] = from_f32<OutT>(t[tx + 32 * j][ty + 8 * i]);
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
static inline int env_int(const char* name, int dflt) { return option(name, dflt); }

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
    p.rois = rois; p.roi_level = nlevels > 1 ? roi_level : nullptr; p.k_dev = nullptr; p.perm = nullptr; p.flags = 0;
    p.C = C; p.K = K; p.PH = PH; p.PW = PW; p.sampling_ratio = sr; p.aligned = aligned;
    return COIN_OK;
}

struct LaunchCfg { int cpl, pwc, threads, chunks, exact; size_t smem; };

static int pick_cfg(LaunchCfg& cfg, int C, int PH, int PW, size_t elt, const char* env_prefix) {
    const int NB = PH * PW;
    cfg.pwc = (PW > 7) ? 14 : 7;
    int cpl = 4;
    char name[64];
    snprintf(name, sizeof name, "%s_CPL", env_prefix);
    cpl = env_int(name, cpl);
    snprintf(name, sizeof name, "%s_PWC", env_prefix);
    cfg.pwc = env_int(name, cfg.pwc);
    if (cfg.pwc != 7 && cfg.pwc != 14) cfg.pwc = 7;
    if (cpl != 1 && cpl != 2 && cpl != 4) cpl = 4;
    while (cpl > 1 && (32 * (cpl / 2) >= C)) cpl /= 2;  // narrow maps: do not waste lanes-by-channel slots
    while (cpl > 1 && (size_t)32 * cpl * NB * elt > 100 * 1024) cpl /= 2;  // keep two CTAs per SM resident
    cfg.cpl = cpl;
    cfg.smem = align_up((size_t)32 * cpl * NB * elt, 128);
    if (cfg.smem > 200 * 1024)
        return fail(COIN_ERR_UNSUPPORTED, "roi_align: output %dx%d needs %zu B of shared memory per CTA", PH, PW, cfg.smem);
    snprintf(name, sizeof name, "%s_ROWS", env_prefix);
    const int rows_per_warp = std::max((int)ceil_div(PH, 8), env_int(name, PH > 7 ? 2 : 1));  // kernels allow <= 8 warps
    const int warps = (int)ceil_div(PH, rows_per_warp);
    cfg.threads = 32 * warps;
    cfg.chunks = (int)ceil_div(C, 32 * cpl);
    cfg.exact = env_int("COIN_ROI_EXACT", 0) == 1;
    return COIN_OK;
}

template <typename OutT, int PWC, int CPL>
static int launch_fwd(const RoiParams& p, OutT* out, const LaunchCfg& cfg, cudaStream_t s) {
    auto kern = cfg.exact ? roi_align_fwd_kernel<OutT, PWC, CPL, true> : roi_align_fwd_kernel<OutT, PWC, CPL, false>;
    if (cfg.smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    kern<<<(unsigned)(p.K * cfg.chunks), cfg.threads, cfg.smem, s>>>(p, out, cfg.chunks);
    return check_launch("roi_align_fwd_kernel");
}

template <typename GT, int PWC, int CPL>
static int launch_bwd(const RoiParams& p, const GT* go, const LaunchCfg& cfg, cudaStream_t s) {
    auto kern = roi_align_bwd_kernel<GT, PWC, CPL>;
    if (cfg.smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    kern<<<(unsigned)(p.K * cfg.chunks), cfg.threads, cfg.smem, s>>>(p, go, cfg.chunks);
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

static int roi_align_fwd_impl(const coin_level_t* levels_host, int nlevels, const float* rois, const int32_t* roi_level,
                              void* out, int out_dtype, int C, int K_cap, int PH, int PW, int sampling_ratio, int aligned,
                              const int32_t* k_dev, const int32_t* perm, coin_stream_t stream) {
    RoiParams p;
    if (int rc = fill_params(p, levels_host, nlevels, rois, roi_level, C, K_cap, PH, PW, sampling_ratio, aligned)) return rc;
    COIN_REQUIRE(out_dtype == COIN_F32 || out_dtype == COIN_F16, "roi_align_fwd: bad out_dtype %d", out_dtype);
    if (K_cap == 0) return COIN_OK;
    COIN_REQUIRE(out, "roi_align_fwd: out is null");
    p.k_dev = k_dev;
    p.perm = perm;
    // COIN_ROI_EXACT: 0 (default) register-tile / separable fast kernels; 1 bit-exact parity kernel; 2 its FMA variant
    const int mode = env_int("COIN_ROI_EXACT", 0);
    if (mode == 0) {
        if (roi_align_fwd_reg_supported(p, out_dtype)) return launch_roi_align_fwd_reg(p, out, out_dtype, as_stream(stream));
        return launch_roi_align_fwd_sep(p, out, out_dtype, as_stream(stream));
    }
    LaunchCfg cfg;
    if (int rc = pick_cfg(cfg, C, PH, PW, out_dtype == COIN_F32 ? 4 : 2, "COIN_ROI_FWD")) return rc;
    cudaStream_t s = as_stream(stream);
    if (out_dtype == COIN_F32) COIN_DISPATCH_ROI(launch_fwd, float, static_cast<float*>(out));
    COIN_DISPATCH_ROI(launch_fwd, __half, static_cast<__half*>(out));
}

extern "C" int coin_roi_align_fwd(const coin_level_t* levels_host, int nlevels, const float* rois,
                                  const int32_t* roi_level, void* out, int out_dtype, int C, int K, int PH,
                                  int PW, int sampling_ratio, int aligned, coin_stream_t stream) {
    return roi_align_fwd_impl(levels_host, nlevels, rois, roi_level, out, out_dtype, C, K, PH, PW, sampling_ratio, aligned,
                              nullptr, nullptr, stream);
}

extern "C" int coin_roi_align_fwd_dev(const coin_level_t* levels_host, int nlevels, const float* rois,
                                      const int32_t* roi_level, void* out, int out_dtype, int C, int K_cap, int PH,
                                      int PW, int sampling_ratio, int aligned, const int32_t* k_dev,
                                      coin_stream_t stream) {
    return roi_align_fwd_impl(levels_host, nlevels, rois, roi_level, out, out_dtype, C, K_cap, PH, PW, sampling_ratio,
                              aligned, k_dev, nullptr, stream);
}

extern "C" int coin_roi_align_fwd_ord(const coin_level_t* levels_host, int nlevels, const float* rois,
                                      const int32_t* roi_level, void* out, int out_dtype, int C, int K_cap, int PH,
                                      int PW, int sampling_ratio, int aligned, const int32_t* k_dev,
                                      const int32_t* perm, coin_stream_t stream) {
    return roi_align_fwd_impl(levels_host, nlevels, rois, roi_level, out, out_dtype, C, K_cap, PH, PW, sampling_ratio,
                              aligned, k_dev, perm, stream);
}

static int roi_align_bwd_impl(const coin_level_t* grad_levels_host, int nlevels, const float* rois,
                              const int32_t* roi_level, const void* grad_out, int grad_dtype, int C, int K,
                              int PH, int PW, int sampling_ratio, int aligned, const int32_t* k_dev, const int32_t* perm,
                              coin_stream_t stream) {
    RoiParams p;
    if (int rc = fill_params(p, grad_levels_host, nlevels, rois, roi_level, C, K, PH, PW, sampling_ratio, aligned)) return rc;
    COIN_REQUIRE(grad_dtype == COIN_F32 || grad_dtype == COIN_F16, "roi_align_bwd: bad grad_dtype %d", grad_dtype);
    if (K == 0) return COIN_OK;
    COIN_REQUIRE(grad_out, "roi_align_bwd: grad_out is null");
    p.perm = perm;
    p.k_dev = k_dev;
    // COIN_ROI_BWD_SEP: 1 (default) the register-tile / separable kernels; 0 the per-sample kernel below
    if (env_int("COIN_ROI_BWD_SEP", 1) != 0 && roi_align_bwd_reg_supported(p, grad_dtype))
        return launch_roi_align_bwd_reg(p, grad_out, grad_dtype, as_stream(stream));
    if (PW <= 32 && env_int("COIN_ROI_BWD_SEP", 1) != 0) return launch_roi_align_bwd_sep(p, grad_out, grad_dtype, as_stream(stream));
    LaunchCfg cfg;
    if (int rc = pick_cfg(cfg, C, PH, PW, grad_dtype == COIN_F32 ? 4 : 2, "COIN_ROI_BWD")) return rc;
    cudaStream_t s = as_stream(stream);
    if (grad_dtype == COIN_F32) COIN_DISPATCH_ROI(launch_bwd, float, static_cast<const float*>(grad_out));
    COIN_DISPATCH_ROI(launch_bwd, __half, static_cast<const __half*>(grad_out));
}

extern "C" int coin_roi_align_bwd(const coin_level_t* grad_levels_host, int nlevels, const float* rois,
                                  const int32_t* roi_level, const void* grad_out, int grad_dtype, int C, int K,
                                  int PH, int PW, int sampling_ratio, int aligned, coin_stream_t stream) {
    return roi_align_bwd_impl(grad_levels_host, nlevels, rois, roi_level, grad_out, grad_dtype, C, K, PH, PW, sampling_ratio,
                              aligned, nullptr, nullptr, stream);
}

extern "C" int coin_roi_align_bwd_ord(const coin_level_t* grad_levels_host, int nlevels, const float* rois,
                                      const int32_t* roi_level, const void* grad_out, int grad_dtype, int C, int K,
                                      int PH, int PW, int sampling_ratio, int aligned, const int32_t* k_dev,
                                      const int32_t* perm, coin_stream_t stream) {
    return roi_align_bwd_impl(grad_levels_host, nlevels, rois, roi_level, grad_out, grad_dtype, C, K, PH, PW, sampling_ratio,
                              aligned, k_dev, perm, stream);
}

// ------------------------------------------------------------------------------------------------
// launch order: the smallest RoIs go last
// ------------------------------------------------------------------------------------------------
// The ROIAlign grids run ~10 waves of CTAs at the reference's batch (1536 RoIs x 4 channel groups on 148 x 4 slots) and a
// RoI's cost grows with its area, so a large RoI that happens to come last leaves most SMs idle for the length of one CTA.
// perm = the input order with the `small_pct` % smallest RoIs (by area) moved to the end and the `big_pct` % largest to the
// front, every part in input order (measured on the bench shape: forward 333 -> 321 us, backward 323 -> 298 us with the
// small ones last; sorting everything by area is slower for the forward, which wants load-heavy and store-heavy CTAs
// mixed). Largest first is for the rare map-sized RoI, whose CTA runs for hundreds of microseconds: it has to start with
// the grid, not wherever the proposal list happens to hold it. One CTA: radix-select of the two area quantiles over
// shared-memory histograms, then a stable three-way partition by a block scan.
constexpr int kOrderThreads = 1024;
constexpr int kOrderMax = 8192;

__global__ void __launch_bounds__(kOrderThreads)
roi_launch_order_kernel(const float* __restrict__ rois, int K_cap, const int32_t* __restrict__ k_dev, int small_pct, int big_pct,
                        int32_t* __restrict__ perm, float area_thr, float side_thr, int divert_cap,
                        int32_t* __restrict__ perm_divert, int32_t* __restrict__ counts) {
    extern __shared__ uint32_t okeys[];     // area bits of every live RoI; bit 0 borrowed: the RoI exceeds the size thresholds
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_rank;
    __shared__ int s_cnt[4][32];
    __shared__ int s_div;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = k_dev ? min(max(__ldg(k_dev), 0), K_cap) : K_cap;
    for (int i = tid; i < K_cap; i += kOrderThreads) {
        if (i < n) {
            const float w = __ldg(rois + 5 * i + 3) - __ldg(rois + 5 * i + 1), h = __ldg(rois + 5 * i + 4) - __ldg(rois + 5 * i + 2);
            const float a = (w > 0.0f && h > 0.0f) ? w * h : 0.0f;     // NaN / inverted: smallest
            const bool div = perm_divert && w > 0.0f && h > 0.0f && (a > area_thr || w > side_thr || h > side_thr);
            // non-negative floats order like their bits (plan mode gives the lowest mantissa bit to the flag)
            okeys[i] = perm_divert ? ((__float_as_uint(a) & ~1u) | (div ? 1u : 0u)) : __float_as_uint(a);
        } else {
            perm[i] = i;
        }
    }
    // RoIs above the size thresholds are DIVERTED: they go to the very end of perm (behind the live count the caller's main
    // launch gets in counts[0]) and into perm_divert, for a launch of their own on the separable kernel - unless there are
    // more of them than that launch's capacity, then nothing is diverted (correct, slow)
    if (tid == 0) s_div = 0;
    __syncthreads();
    if (perm_divert) {
        int nd = 0;
        for (int i = tid; i < n; i += kOrderThreads) nd += okeys[i] & 1u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, o);
        if (lane == 0 && nd) atomicAdd(&s_div, nd);
    }
    __syncthreads();
    const bool divert_on = perm_divert && s_div > 0 && s_div <= divert_cap;
    // the key of a given rank (0-based, ascending), one byte per pass over shared-memory histograms
    auto select = [&](int rank_wanted) -> uint32_t {
        __syncthreads();
        if (tid == 0) { s_prefix = 0; s_rank = (uint32_t)rank_wanted; }
        for (int pass = 3; pass >= 0; --pass) {
            const int shift = 8 * pass;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < n; i += kOrderThreads) {
                const uint32_t k = okeys[i];
                if (pass == 3 || (k >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t rank = s_rank, b = 0;
                while (b < 255 && hist[b] <= rank) { rank -= hist[b]; ++b; }
                s_rank = rank;
                s_prefix = prefix | (b << shift);
            }
            __syncthreads();
        }
        return s_prefix;
    };
    const int n_small = (int)((long long)n * small_pct / 100), n_big = (int)((long long)n * big_pct / 100);
    // RoIs with key <= lo are "small" (launched last), with key >= hi "big" (launched first); ties may add a few
    const uint32_t lo = n_small > 0 ? select(n_small - 1) : 0u;
    const bool any_small = n_small > 0;
    const uint32_t hi = n_big > 0 ? select(n - n_big) : 0xffffffffu;
    const bool any_big = n_big > 0;
    auto cls = [&](uint32_t k) {
        if (divert_on && (k & 1u)) return 3;      // (divert_on implies plan mode: bit 0 is the flag)
        return (any_big && k >= hi && !(any_small && k <= lo)) ? 0 : ((any_small && k <= lo) ? 2 : 1);
    };
    const int per = (n + kOrderThreads - 1) / kOrderThreads;
    const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
    int cnt[4] = {0, 0, 0, 0};
    for (int i = i0; i < i1; ++i) ++cnt[cls(okeys[i])];
    int pre[4] = {cnt[0], cnt[1], cnt[2], cnt[3]};     // inclusive warp scans
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int v = __shfl_up_sync(0xffffffffu, pre[c], o);
            if (lane >= o) pre[c] += v;
        }
    if (lane == 31) { s_cnt[0][warp] = pre[0]; s_cnt[1][warp] = pre[1]; s_cnt[2][warp] = pre[2]; s_cnt[3][warp] = pre[3]; }
    __syncthreads();
    int off[4], tot[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c) off[c] = pre[c] - cnt[c];
    for (int w = 0; w < kOrderThreads / 32; ++w)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (w < warp) off[c] += s_cnt[c][w];
            tot[c] += s_cnt[c][w];
        }
    const int nd0 = off[3];                   // this thread's first index in the diverted list
    off[1] += tot[0];
    off[2] += tot[0] + tot[1];
    off[3] += tot[0] + tot[1] + tot[2];
    int jd = nd0;
    for (int i = i0; i < i1; ++i) {
        const int c = cls(okeys[i]);
        perm[off[c]++] = i;
        if (c == 3) perm_divert[jd++] = i;
    }
    if (tid == 0 && counts) { counts[0] = n - tot[3]; counts[1] = tot[3]; }
}

// Two RoI lists by box size (in input order, or in the order of a launch order `order`): "big" = area > area_thr or a side > side_thr, and the rest, with their
// device-side lengths. The step pools the few map-sized or map-wide private boxes with the separable kernel and the rest
// with the register-tile kernel (a register-tile CTA walks such a box for 0.2 - 1 ms: tools/c_box_probe.py).
__global__ void __launch_bounds__(kOrderThreads)
roi_split_by_area_kernel(const float* __restrict__ rois, int K_cap, const int32_t* __restrict__ k_dev, float area_thr,
                         float side_thr, int big_cap, const int32_t* __restrict__ order, int32_t* __restrict__ perm_small, int32_t* __restrict__ perm_big, int32_t* __restrict__ counts) {
    __shared__ int s_cnt[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = k_dev ? min(max(__ldg(k_dev), 0), K_cap) : K_cap;
    auto is_big = [&](int pos) {
        const int i = order ? order[pos] : pos;
        const float w = __ldg(rois + 5 * i + 3) - __ldg(rois + 5 * i + 1), h = __ldg(rois + 5 * i + 4) - __ldg(rois + 5 * i + 2);
        return w > 0.0f && h > 0.0f && (w * h > area_thr || w > side_thr || h > side_thr);
    };
    const int per = (n + kOrderThreads - 1) / kOrderThreads;
    const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
    int nb = 0;
    for (int i = i0; i < i1; ++i) nb += is_big(i);
    int pb = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int vb = __shfl_up_sync(0xffffffffu, pb, o);
        if (lane >= o) pb += vb;
    }
    if (lane == 31) s_cnt[1][warp] = pb;
    __syncthreads();
    int ob = pb - nb, tbig = 0;
    for (int w = 0; w < kOrderThreads / 32; ++w) {
        if (w < warp) ob += s_cnt[1][w];
        tbig += s_cnt[1][w];
    }
    // the big list holds at most big_cap RoIs (its launch is sized for that); big RoIs beyond it stay in the other list
    for (int i = i0; i < i1; ++i) {
        const int idx = order ? order[i] : i;
        if (is_big(i)) {
            if (ob < big_cap) perm_big[ob] = idx; else perm_small[i - big_cap] = idx;
            ++ob;
        } else {
            perm_small[i - min(ob, big_cap)] = idx;
        }
    }
    if (tid == 0) { counts[1] = min(tbig, big_cap); counts[0] = n - min(tbig, big_cap); }
}

extern "C" int coin_roi_split_by_area(const float* rois, int K_cap, const int32_t* k_dev, float area_thr, float side_thr,
                                      int big_cap, const int32_t* order, int32_t* perm_small,
                                      int32_t* perm_big, int32_t* counts, coin_stream_t stream) {
    COIN_REQUIRE(K_cap >= 0 && big_cap >= 0 && counts, "roi_split_by_area: bad arguments");
    cudaStream_t s = as_stream(stream);
    if (K_cap == 0) { fill_bytes(counts, 0, 2 * sizeof(int32_t), s); return COIN_OK; }
    COIN_REQUIRE(rois && perm_small && perm_big, "roi_split_by_area: null pointer");
    roi_split_by_area_kernel<<<1, kOrderThreads, 0, s>>>(rois, K_cap, k_dev, area_thr, side_thr, big_cap, order, perm_small, perm_big, counts);
    return check_launch("roi_split_by_area_kernel");
}

static int roi_launch_order_impl(const float* rois, int K_cap, const int32_t* k_dev, int small_pct, int big_pct, int32_t* perm,
                                 float area_thr, float side_thr, int divert_cap, int32_t* perm_divert, int32_t* counts,
                                 coin_stream_t stream) {
    COIN_REQUIRE(K_cap >= 0 && small_pct >= 0 && big_pct >= 0 && small_pct + big_pct <= 100, "roi_launch_order: bad arguments");
    if (K_cap == 0) {
        if (counts) fill_bytes(counts, 0, 2 * sizeof(int32_t), as_stream(stream));
        return COIN_OK;
    }
    COIN_REQUIRE(rois && perm, "roi_launch_order: null pointer");
    COIN_REQUIRE(K_cap <= kOrderMax, "roi_launch_order: K=%d exceeds %d (large grids have no tail worth ordering)", K_cap, kOrderMax);
    roi_launch_order_kernel<<<1, kOrderThreads, (size_t)K_cap * sizeof(uint32_t), as_stream(stream)>>>(
        rois, K_cap, k_dev, small_pct, big_pct, perm, area_thr, side_thr, divert_cap, perm_divert, counts);
    return check_launch("roi_launch_order_kernel");
}

extern "C" int coin_roi_launch_order(const float* rois, int K_cap, const int32_t* k_dev, int small_pct, int big_pct,
                                     int32_t* perm, coin_stream_t stream) {
    return roi_launch_order_impl(rois, K_cap, k_dev, small_pct, big_pct, perm, 0.0f, 0.0f, 0, nullptr, nullptr, stream);
}

extern "C" int coin_roi_launch_plan(const float* rois, int K_cap, const int32_t* k_dev, int small_pct, int big_pct,
                                    float area_thr, float side_thr, int divert_cap, int32_t* perm, int32_t* perm_divert,
                                    int32_t* counts, coin_stream_t stream) {
    COIN_REQUIRE(perm_divert && counts && divert_cap >= 0, "roi_launch_plan: null pointer");
    return roi_launch_order_impl(rois, K_cap, k_dev, small_pct, big_pct, perm, area_thr, side_thr, divert_cap, perm_divert,
                                 counts, stream);
}

extern "C" int coin_nchw_to_nhwc_f32(const void* in, int in_dtype, float* out, int N, int C, int H, int W,
                                     coin_stream_t stream) {
    COIN_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad shape");
    COIN_REQUIRE(in_dtype == COIN_F32 || in_dtype == COIN_F16, "nchw_to_nhwc: bad dtype %d", in_dtype);
    if (N == 0) return COIN_OK;
    COIN_REQUIRE(in && out, "nchw_to_nhwc: null pointer");
    const int HW = H * W;
    dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(C, 32), (unsigned)N), block(32, 8);
    if (env_int("COIN_LAYOUT_V2", 1)) {
        dim3 g2((unsigned)ceil_div(HW, kLtHw), (unsigned)ceil_div(C, 32), (unsigned)N);
        if (in_dtype == COIN_F32)
            nchw_to_nhwc_v2_kernel<float><<<g2, block, 0, as_stream(stream)>>>(static_cast<const float*>(in), out, C, HW);
        else
            nchw_to_nhwc_v2_kernel<__half><<<g2, block, 0, as_stream(stream)>>>(static_cast<const __half*>(in), out, C, HW);
        return check_launch("nchw_to_nhwc_v2_kernel");
    }
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
    if (env_int("COIN_LAYOUT_V2", 1)) {
        dim3 g2((unsigned)ceil_div(HW, kLtHw), (unsigned)ceil_div(C, 32), (unsigned)N);
        if (out_dtype == COIN_F32)
            nhwc_to_nchw_v2_kernel<float><<<g2, block, 0, as_stream(stream)>>>(in, static_cast<float*>(out), C, HW);
        else
            nhwc_to_nchw_v2_kernel<__half><<<g2, block, 0, as_stream(stream)>>>(in, static_cast<__half*>(out), C, HW);
        return check_launch("nhwc_to_nchw_v2_kernel");
    }
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
    COIN_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "roi_pooler_levels: boxes must be 16-byte aligned");
    pooler_levels_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        boxes, n, min_level, max_level, (float)canonical_box_size, (float)canonical_level, out_levels);
    return check_launch("pooler_levels_kernel");
}
