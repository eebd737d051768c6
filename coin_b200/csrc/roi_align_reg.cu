// roi_align_reg.cu -- register-tile ROIAlign forward / backward for sm_100a (the default kernels for the
// reference's head shapes: 14x14 and 7x7 outputs, C a multiple of 32, fp32).
//
// Replaces torchvision::roi_align / torchvision::_roi_align_backward as reached from
// coin/modeling/roi_heads/clip_roi_heads.py:51-63,142-147,172-176 (ROIPooler -> ROIAlign). Same sample
// positions, bilinear weights, validity rule and 1/count scaling as the torchvision kernels; the summation is
// re-associated through the separable form (roi_align_sep.cu), so results agree to ~1e-7 relative.
//
// What changed against roi_align_sep.cu, and why (profiles/r01g_roi_align_ncu.md): that kernel spent one
// shared-memory load per FMA (lane = output bin walking channels) and ~1430 LSU wavefronts per RoI and 32
// channels, which pinned the L1/LSU data pipe at ~70 % while DRAM idled at 33 %. Here a lane IS a channel for
// the whole computation and the unit's R x PW outputs live in registers:
//   forward   for every feature column of the RoI: t_r = sum_y Wy[r][y] * F[y][x][c]   (coalesced 128-byte loads,
//             the y taps of the unit's R rows merged into one table so shared feature rows are loaded once),
//             then acc[r][pw] += Wx[x][pw] * t_r for the <= 4 bins that touch the column (warp-uniform window
//             start -> a switch over statically indexed registers; RoIs narrower than ~7 cells use all PW bins).
//             The accumulators go to a [32 channels][PH*PW] tile in shared memory laid out exactly like the
//             CTA's contiguous output region out[k, c0:c0+32, :, :], and the tile leaves through ONE 1-D TMA
//             bulk store (25 KB for 14x14) with an L2 evict-first hint - no LDS / STG for the 1.2 GB output.
//   backward  the mirror image: the [32][PH*PW] grad_out tile arrives through ONE TMA bulk load, each lane
//             pulls its channel's R x PW values into registers, u_r = sum_pw Wx[x][pw] * g[r][pw] per column,
//             and one fp32 RED per touched (feature row, column, channel) flushes sum_r Wy[r][y] * u_r.
// ~480 LSU wavefronts per RoI and 32 channels instead of ~1430 / ~1200.
#include <climits>

#include "roi_common.cuh"

namespace coin {

static inline int reg_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

constexpr int kRegTap = 128;    // x / y tap-table entries (PW*grid_w and PH*grid_h must fit)
constexpr int kRegCols = 96;    // feature columns of one RoI handled by the tables
constexpr int kRegYEnt = 16;    // merged y-table entries per unit

struct RXTap { int lo, hi; float l, h; };    // cell indices; lo < 0: sample outside the map

__device__ __forceinline__ RXTap reg_xtap(float start, float bin, int p, int i, int grid, int size) {
    const float v = start + (float)p * bin + ((float)i + 0.5f) * bin / (float)grid;
    RXTap t;
    if (!axis_taps(v, size, t.lo, t.hi, t.l, t.h)) { t.lo = -1; t.hi = -1; t.l = 0.0f; t.h = 0.0f; }
    return t;
}

__device__ __forceinline__ uint32_t reg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void reg_bulk_store(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(reg_smem_u32(ssrc)), "r"(bytes), "l"(pol)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void reg_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void reg_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Packed fp32 FMA of Blackwell (FFMA2): d = a * b + c on register pairs, b broadcast to both halves (the
// compiler folds the {b, b} pair into the instruction's scalar-operand form: no MOV).
__device__ __forceinline__ float2 ffma2(const float2 a, const float b, const float2 c) {
    unsigned long long xa, xb, xc, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(xa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(xb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(xc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(xa), "l"(xb), "l"(xc));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(d));
    return r;
}

struct __align__(16) ColWin { float w[6]; int p0h; int pad; };   // the <= 6 non-zero weights of a column, bins [2*p0h, 2*p0h + 6)

// acc[r][H + i] += w[2i .. 2i+1] * t[r] for the 6-bin window starting at the compile-time bin pair H
template <int R, int NP, int H>
__device__ __forceinline__ void win6_apply(float2 (&acc)[R][NP], const float4 wa, const float2 wb, const float (&t)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        acc[r][H + 0] = ffma2(make_float2(wa.x, wa.y), t[r], acc[r][H + 0]);
        acc[r][H + 1] = ffma2(make_float2(wa.z, wa.w), t[r], acc[r][H + 1]);
        acc[r][H + 2] = ffma2(wb, t[r], acc[r][H + 2]);
    }
}

// the window start is warp-uniform: a jump table over statically indexed accumulators
template <int R, int NP>
__device__ __forceinline__ void win6_switch(float2 (&acc)[R][NP], const int p0h, const float4 wa, const float2 wb,
                                            const float (&t)[R]) {
    static_assert(NP >= 3 && NP <= 8, "window switch covers 6 <= PW <= 16");
    switch (p0h) {
#define COIN_WIN_CASE(H) \
    case H:              \
        if constexpr (H + 3 <= NP) win6_apply<R, NP, (H + 3 <= NP ? H : 0)>(acc, wa, wb, t); \
        break;
        COIN_WIN_CASE(0) COIN_WIN_CASE(1) COIN_WIN_CASE(2) COIN_WIN_CASE(3) COIN_WIN_CASE(4) COIN_WIN_CASE(5)
#undef COIN_WIN_CASE
        default: break;
    }
}

// shared tables of one RoI (built once per CTA); units are pairs of output rows
template <int NU>
struct RegTables {
    float wxd[kRegCols + 2][16];          // combined x weight (already / count) of every bin on every column
    ColWin win[kRegCols + 2];             // the same weights as a 6-bin window (valid when maxspan <= 6)
    int yoff[NU][kRegYEnt];               // merged y table of every unit: feature-row offset (y*W*C) ...
    float2 yw[NU][kRegYEnt];              // ... and its weights on the unit's two output rows
    int ycnt[NU];
    int mode, cmin, cmax, maxspan;        // mode 0: tables, 1: direct evaluation, 2: the RoI pools to zeros
};

// Builds the tables; every thread of the CTA must call it. `scratch` (>= 2*kRegTap*16 bytes, 16-byte aligned) holds
// the per-sample tap tables while the merged tables are built and is free again on return.
template <int PH, int PW, int NT>
__device__ __forceinline__ void reg_build_tables(RegTables<(PH + 1) / 2>& tb, void* scratch, const RoiGeom& g,
                                                 const int H, const int W, const int C) {
    constexpr int NU = (PH + 1) / 2;
    RXTap* xs = reinterpret_cast<RXTap*>(scratch);
    Tap* ys = reinterpret_cast<Tap*>(scratch) + kRegTap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gh = g.grid_h, gw = g.grid_w;
    const bool empty = gh <= 0 || gw <= 0;
    const bool tables = !empty && (long long)PW * gw <= kRegTap && (long long)PH * gh <= kRegTap;
    const float rcount = 1.0f / g.count;
    if (tid == 0) { tb.mode = empty ? 2 : (tables ? 0 : 1); tb.cmin = INT_MAX; tb.cmax = -1; tb.maxspan = 0; }
    if (tables) {
        const int nx = PW * gw, ny = PH * gh;
        for (int s = tid; s < nx; s += NT) {
            const int pw = s / gw;
            xs[s] = reg_xtap(g.start_w, g.bin_w, pw, s - pw * gw, gw, W);
        }
        for (int s = NT - 1 - tid; s < ny; s += NT) {
            const int ph = s / gh;
            ys[s] = make_tap(g.start_h, g.bin_h, ph, s - ph * gh, gh, H, W * C);
        }
    }
    __syncthreads();
    if (tables) {
        for (int u = NT - 1 - tid; u < NU; u += NT) {   // merged y table: one thread per unit
            int n = 0;
            bool overflow = false;
            for (int r = 0; r < 2 && u * 2 + r < PH && !overflow; ++r)
                for (int iy = 0; iy < gh && !overflow; ++iy) {
                    const Tap Y = ys[(u * 2 + r) * gh + iy];
                    if (Y.lo < 0) continue;
                    for (int t = 0; t < 2; ++t) {
                        const int off = t ? Y.hi : Y.lo;
                        const float w = t ? Y.l : Y.h;
                        int e = 0;
                        while (e < n && tb.yoff[u][e] != off) ++e;
                        if (e == n) {
                            if (n == kRegYEnt) { overflow = true; break; }
                            tb.yoff[u][e] = off;
                            tb.yw[u][e] = make_float2(0.0f, 0.0f);
                            ++n;
                        }
                        if (r == 0) tb.yw[u][e].x += w; else tb.yw[u][e].y += w;
                    }
                }
            tb.ycnt[u] = n;
            if (overflow) tb.mode = 1;
        }
        if (warp == 0) {   // feature-column range of the RoI
            int lo = INT_MAX, hi = -1;
            for (int s = lane; s < PW * gw; s += 32) {
                const RXTap X = xs[s];
                if (X.lo >= 0) { lo = min(lo, X.lo); hi = max(hi, X.hi); }
            }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0) {
                tb.cmin = lo; tb.cmax = hi;
                if (hi < 0) tb.mode = 2;                          // every x sample lies outside the map
                else if (hi - lo + 1 > kRegCols) tb.mode = 1;
            }
        }
    }
    __syncthreads();
    if (tb.mode != 0) return;
    const int cmin = tb.cmin, ncols = tb.cmax - cmin + 1;
    for (int idx = tid; idx < (ncols + 2) * 16; idx += NT) {   // dense weights (+ two all-zero padding columns)
        const int ci = idx >> 4, pw = idx & 15, col = cmin + ci;
        float w = 0.0f;
        if (pw < PW && ci < ncols)
            for (int ix = 0; ix < gw; ++ix) {
                const RXTap X = xs[pw * gw + ix];
                if (X.lo < 0) continue;
                if (X.lo == col) w += X.h;
                if (X.hi == col) w += X.l;
            }
        tb.wxd[ci][pw] = w * rcount;
    }
    __syncthreads();
    for (int ci = tid; ci < ncols + 2; ci += NT) {             // window of non-zero bins per column
        int first = -1, last = -1;
#pragma unroll
        for (int pw = 0; pw < PW; ++pw)
            if (tb.wxd[ci][pw] != 0.0f) { if (first < 0) first = pw; last = pw; }
        constexpr int kMaxH = (PW + 1) / 2 >= 3 ? (PW + 1) / 2 - 3 : 0;      // last window start (in bin pairs)
        const int p0h = first < 0 ? 0 : min(first >> 1, kMaxH);
        ColWin cw;
#pragma unroll
        for (int i = 0; i < 6; ++i) cw.w[i] = tb.wxd[ci][2 * p0h + i];       // columns >= PW of wxd are zero
        cw.p0h = p0h; cw.pad = 0;
        tb.win[ci] = cw;
        if (first >= 0) atomicMax(&tb.maxspan, last - 2 * p0h + 1);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// x phase of one feature column: acc[r][pw] += Wx[col][pw] * t[r]
template <int PW, bool DENSE, typename TB>
__device__ __forceinline__ void fwd_xphase(float2 (&acc)[2][(PW + 1) / 2], const TB& tb, const int ci, const float (&t)[2]) {
    constexpr int NP = (PW + 1) / 2;
    if (DENSE) {
        const float4* wrow = reinterpret_cast<const float4*>(tb.wxd[ci]);
#pragma unroll
        for (int i = 0; i < (NP + 1) / 2; ++i) {
            const float4 x = wrow[i];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                acc[r][2 * i] = ffma2(make_float2(x.x, x.y), t[r], acc[r][2 * i]);
                if (2 * i + 1 < NP) acc[r][2 * i + 1] = ffma2(make_float2(x.z, x.w), t[r], acc[r][2 * i + 1]);
            }
        }
    } else {
        const float4* cw = reinterpret_cast<const float4*>(&tb.win[ci]);
        const float4 wa = cw[0], wb = cw[1];
        win6_switch<2, NP>(acc, __float_as_int(wb.z), wa, make_float2(wb.x, wb.y), t);
    }
}

// NQ feature columns at pj[.] (+ q * cstride): all NJ * NQ loads first, then the y and x phases
template <int PW, int CS, bool DENSE, int NJ, int NQ, typename TB>
__device__ __forceinline__ void fwd_step(float2 (&acc)[2][(PW + 1) / 2], const TB& tb, const int ci,
                                         const float* const (&pj)[NJ], const float2 (&yw)[NJ], const int C) {
    const int cstride = CS ? CS : C;
    float v[NQ][NJ];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < NJ; ++j) v[q][j] = __ldg(pj[j] + q * cstride);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float2 t2 = make_float2(yw[0].x * v[q][0], yw[0].y * v[q][0]);
#pragma unroll
        for (int j = 1; j < NJ; ++j) t2 = ffma2(yw[j], v[q][j], t2);
        const float t[2] = {t2.x, t2.y};
        fwd_xphase<PW, DENSE>(acc, tb, ci + q, t);
    }
}

// One pass over the RoI's feature columns for one unit and NJ merged y entries starting at e0.
template <int PW, int CS, bool DENSE, int NJ, typename TB>
__device__ __forceinline__ void fwd_columns(float2 (&acc)[2][(PW + 1) / 2], const TB& tb, const int u, const int e0,
                                            const float* __restrict__ fcol, const int ncols, const int C) {
    const float* pj[NJ];
    float2 yw[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        pj[j] = fcol + tb.yoff[u][e0 + j];
        yw[j] = tb.yw[u][e0 + j];
    }
    const int cstride = CS ? CS : C;
    constexpr int NQ = NJ <= 2 ? 4 : 2;       // columns per step: 8 independent loads in flight
    int ci = 0;
    for (; ci + NQ <= ncols; ci += NQ) {
        fwd_step<PW, CS, DENSE, NJ, NQ>(acc, tb, ci, pj, yw, C);
#pragma unroll
        for (int j = 0; j < NJ; ++j) pj[j] += NQ * cstride;
    }
    for (; ci < ncols; ++ci) {
        fwd_step<PW, CS, DENSE, NJ, 1>(acc, tb, ci, pj, yw, C);
#pragma unroll
        for (int j = 0; j < NJ; ++j) pj[j] += cstride;
    }
}

template <int PW, int CS, bool DENSE, typename TB>
__device__ __forceinline__ void fwd_unit(float2 (&acc)[2][(PW + 1) / 2], const TB& tb, const int u, const int ne,
                                         const float* __restrict__ fcol, const int ncols, const int C) {
    for (int e0 = 0; e0 < ne; e0 += 4) {
        switch (min(4, ne - e0)) {   // warp-uniform
            case 1: fwd_columns<PW, CS, DENSE, 1>(acc, tb, u, e0, fcol, ncols, C); break;
            case 2: fwd_columns<PW, CS, DENSE, 2>(acc, tb, u, e0, fcol, ncols, C); break;
            case 3: fwd_columns<PW, CS, DENSE, 3>(acc, tb, u, e0, fcol, ncols, C); break;
            default: fwd_columns<PW, CS, DENSE, 4>(acc, tb, u, e0, fcol, ncols, C); break;
        }
    }
}

template <int PH, int PW, int CS>
__global__ void __launch_bounds__(32 * ((PH + 1) / 2), (PH * PW > 64 ? 4 : 8))
roi_align_fwd_reg_kernel(const RoiParams p, float* __restrict__ out, const int cgroups, const int slabs) {
    constexpr int NU = (PH + 1) / 2, NB = PH * PW, NT = 32 * NU, NP = (PW + 1) / 2;
    constexpr bool VEC = PW % 2 == 0 && (2 * PW) % 4 == 0 && NB % 4 == 0;
    extern __shared__ __align__(128) float tile[];   // [32 channels][NB]: the CTA's contiguous output region
    __shared__ RegTables<NU> tb;
    static_assert(32 * NB * sizeof(float) >= 2 * kRegTap * 16, "the tile doubles as tap-table scratch");

    const int k = blockIdx.x / cgroups;
    if (p.k_dev && k >= __ldg(p.k_dev)) return;   // capacity launch: RoI beyond the live count
    const int cg0 = (blockIdx.x - k * cgroups) * (32 * slabs);
    const int lane = threadIdx.x & 31, u = threadIdx.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W;
    const int C = CS ? CS : p.C;
    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    const float* __restrict__ fimg = L.feat_nhwc + (size_t)g.batch * H * W * C;
    const int nslab = min(slabs, (C - cg0) / 32);
    float* __restrict__ oroi = out + (size_t)k * C * NB;

    reg_build_tables<PH, PW, NT>(tb, tile, g, H, W, C);
    const int mode = tb.mode;

    if (mode == 2) {   // no sample inside the map: the RoI pools to zeros
        float* o = oroi + (size_t)cg0 * NB;
        for (int e = threadIdx.x; e < nslab * 32 * NB; e += NT) o[e] = 0.0f;
        return;
    }
    if (mode == 1) {   // exotic geometry (sampling grids beyond the tables): direct 4-tap evaluation
        const float rcount = 1.0f / g.count;
        for (int sl = 0; sl < nslab; ++sl) {
            const int c = cg0 + sl * 32 + lane;
            for (int b = u; b < NB; b += NU) {
                const int ph = b / PW, pw = b - ph * PW;
                float acc = 0.0f;
                for (int iy = 0; iy < g.grid_h; ++iy) {
                    const Tap Y = make_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H, W * C);
                    if (Y.lo < 0) continue;
                    for (int ix = 0; ix < g.grid_w; ++ix) {
                        const Tap X = make_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W, C);
                        if (X.lo < 0) continue;
                        const float* f = fimg + c;
                        acc += Y.h * X.h * __ldg(f + Y.lo + X.lo) + Y.h * X.l * __ldg(f + Y.lo + X.hi) +
                               Y.l * X.h * __ldg(f + Y.hi + X.lo) + Y.l * X.l * __ldg(f + Y.hi + X.hi);
                    }
                }
                oroi[(size_t)c * NB + b] = acc * rcount;
            }
        }
        return;
    }

    const int cmin = tb.cmin, ncols = tb.cmax - cmin + 1;
    const bool dense = PW < 6 || tb.maxspan > 6;
    const int ne = tb.ycnt[u];
    const int nr = min(2, PH - u * 2);
    const uint64_t pol = l2_evict_first_policy();

    for (int sl = 0; sl < nslab; ++sl) {
        const int c0 = cg0 + sl * 32;
        const float* __restrict__ fcol = fimg + c0 + lane + (size_t)cmin * C;
        float2 acc[2][NP];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < NP; ++i) acc[r][i] = make_float2(0.0f, 0.0f);
        if (dense) fwd_unit<PW, CS, true>(acc, tb, u, ne, fcol, ncols, C);
        else fwd_unit<PW, CS, false>(acc, tb, u, ne, fcol, ncols, C);
        // the previous slab's bulk store must have read the tile before it is overwritten
        if (sl > 0 && threadIdx.x == 0) reg_bulk_wait_read();
        __syncthreads();
        float* __restrict__ tp = tile + lane * NB + u * (2 * PW);
        if (VEC) {   // rows 2u, 2u+1 are 2*PW consecutive floats of the channel's plane
            float4* tp4 = reinterpret_cast<float4*>(tp);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const int r = (2 * i) / NP, a = 2 * i - r * NP;           // pair index 2i, 2i+1 of the 2*NP pairs
                const int r2 = (2 * i + 1) / NP, b = 2 * i + 1 - r2 * NP;
                tp4[i] = make_float4(acc[r][a].x, acc[r][a].y, acc[r2][b].x, acc[r2][b].y);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int b = 0; b < PW; ++b)
                    if (r < nr) tp[r * PW + b] = (b & 1) ? acc[r][b >> 1].y : acc[r][b >> 1].x;
        }
        reg_fence_async();
        __syncthreads();
        if (threadIdx.x == 0) reg_bulk_store(oroi + (size_t)c0 * NB, tile, 32 * NB * sizeof(float), pol);
    }
    if (threadIdx.x == 0) reg_bulk_wait_read();
}

template <int PH, int PW, int CS>
static int launch_fwd_reg(const RoiParams& p, float* out, int slabs, cudaStream_t s) {
    constexpr int NU = (PH + 1) / 2, NB = PH * PW;
    auto kern = roi_align_fwd_reg_kernel<PH, PW, CS>;
    const size_t smem = (size_t)32 * NB * sizeof(float);
    const int cgroups = (int)ceil_div(p.C, 32 * slabs);
    kern<<<(unsigned)(p.K * cgroups), 32 * NU, smem, s>>>(p, out, cgroups, slabs);
    return check_launch("roi_align_fwd_reg_kernel");
}

bool roi_align_fwd_reg_supported(const RoiParams& p, int out_dtype) {
    if (reg_env("COIN_ROI_REG", 1) == 0) return false;
    if (out_dtype != COIN_F32 || p.C % 32 != 0) return false;
    return (p.PH == 14 && p.PW == 14) || (p.PH == 7 && p.PW == 7);
}

int launch_roi_align_fwd_reg(const RoiParams& p, void* out, cudaStream_t s) {
    if (reinterpret_cast<uintptr_t>(out) & 15) return fail(COIN_ERR_INVALID, "roi_align_fwd: out must be 16-byte aligned");
    float* o = static_cast<float*>(out);
    const int nsl = (int)(p.C / 32);
    // few RoIs: fewer channel slabs per CTA so that one very large RoI cannot leave a long tail
    int slabs = reg_env("COIN_ROI_REG_SLABS", p.K < 1024 ? 2 : 8);
    slabs = std::max(1, std::min(slabs, nsl));
    if (p.PH == 14) {
        if (p.C == 1024) return launch_fwd_reg<14, 14, 1024>(p, o, slabs, s);
        return launch_fwd_reg<14, 14, 0>(p, o, slabs, s);
    }
    if (p.C == 1024) return launch_fwd_reg<7, 7, 1024>(p, o, slabs, s);
    return launch_fwd_reg<7, 7, 0>(p, o, slabs, s);
}

}  // namespace coin
