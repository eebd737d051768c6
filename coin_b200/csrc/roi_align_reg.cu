// roi_align_reg.cu -- register-tile ROIAlign forward / backward for sm_100a (the default kernels for the
// reference's head shapes: 14x14 and 7x7 outputs, C a multiple of 32; fp32 or fp16 I/O, fp32 arithmetic).
//
// Replaces torchvision::roi_align / torchvision::_roi_align_backward as reached from
// coin/modeling/roi_heads/clip_roi_heads.py:51-63,142-147,172-176 (ROIPooler -> ROIAlign). Same sample
// positions, bilinear weights, validity rule and 1/count scaling as the torchvision kernels; the summation is
// re-associated through the separable form (roi_align_sep.cu), so results agree to ~1e-7 relative.
//
// What changed against roi_align_sep.cu, and why (profiles/r01g_roi_align_ncu.md, r01h_roi_align_ncu.md): that
// kernel spent one shared-memory load per FMA (lane = output bin walking channels) and ~1430 LSU wavefronts per
// RoI and 32 channels, which pinned the L1/LSU data pipe at ~70 % while DRAM idled at 33 %. Here a lane IS a
// channel for the whole computation and the unit's 2 x PW outputs live in registers as FFMA2 pairs:
//   forward   for every feature column of the RoI: t_r = sum_y Wy[r][y] * F[y][x][c]   (coalesced 128-byte loads,
//             the y taps of the unit's two rows merged into one table so shared feature rows are loaded once),
//             then acc[r][pw] += Wx[x][pw] * t_r with the column's weight row broadcast from shared memory. The
//             bins are walked in two groups (left / right half of the row) with their own column ranges, and a
//             step covers as many columns as keep 8-12 independent loads in flight per warp.
//             The accumulators go to a [32 channels][PH*PW] tile in shared memory laid out exactly like the
//             CTA's contiguous output region out[k, c0:c0+32, :, :], and the tile leaves through ONE 1-D TMA
//             bulk store (25 KB for 14x14) with an L2 evict-first hint - no LDS / STG for the 1.2 GB output.
//   backward  the mirror image: the [32][PH*PW] grad_out tile arrives through ONE TMA bulk load, each lane
//             pulls its channel's 2 x PW values into registers, u_r = sum_pw Wx[x][pw] * g[r][pw] per column,
//             and one fp32 RED per touched (feature row, column, channel) flushes sum_r Wy[r][y] * u_r.
#include <climits>

#include "roi_common.cuh"

namespace coin {

static inline int reg_env(const char* name, int dflt) { return option(name, dflt); }

// COIN_ROI_CTAS_PER_SM = n > 0: pad the dynamic shared memory of the register-tile kernels so that at most n CTAs fit on
// an SM. The step (coin_b200/pipeline.py) uses it to leave a quarter of every SM to the latency-bound kernels that run
// beside ROIAlign on other streams: with 4 resident CTAs of 224 threads x 72 registers the register file is full, and a
// short kernel's CTA has to wait for a ROIAlign CTA (~30 us) to retire on an SM with enough room.
static inline size_t reg_occupancy_pad(size_t smem) {
    const int occ = reg_env("COIN_ROI_CTAS_PER_SM", 0);
    if (occ <= 0) return smem;
    const size_t per_cta = ((size_t)228 * 1024) / (size_t)occ - 1024;   // 1 KB per CTA is reserved by the runtime
    return std::max(smem, per_cta / 128 * 128);
}

// COIN_ROI_CARVEOUT = p in [0, 100]: preferred shared-memory carveout (percent of the unified L1 / shared array) of the
// register-tile kernels. Left to the driver, the carveout is sized for the 4 resident ROIAlign CTAs (~138 KB), and a
// kernel of another stream that needs > ~26 KB of shared memory (the single-CTA sort, knowledge separation) cannot
// become resident on any SM until the ROIAlign grid has drained. The step asks for the maximum carveout instead.
template <class K>
static inline void reg_apply_carveout(K kern) {
    const int pct = reg_env("COIN_ROI_CARVEOUT", -1);
    if (pct >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
}

constexpr int kRegTap = 256;    // x / y tap-table entries (PW*grid_w and PH*grid_h must fit: RoIs up to 18 cells per bin)
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

__device__ __forceinline__ void reg_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(reg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void reg_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(reg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void reg_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(
            reg_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

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

__device__ __forceinline__ float2 ffma2v(const float2 a, const float2 b, const float2 c) {   // a * b + c, pairwise
    unsigned long long xa, xb, xc, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(xa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(xb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(xc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(xa), "l"(xb), "l"(xc));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(d));
    return r;
}

// shared tables of one RoI (built once per CTA); units are pairs of output rows
template <int NU>
struct RegTables {
    float wxd[kRegCols + 2][16];          // combined x weight (already / count) of every bin on every column
    float2 yw[NU][kRegYEnt];              // merged y table of a unit: weights of feature row ymin[u] + e on its two rows
    int ymin[NU], ycnt[NU];
    int mode;                             // 0: tables, 1: direct evaluation, 2: the RoI pools to zeros
};

// Builds the tables (two CTA barriers); every thread of the CTA must call it. `scratch` (>= 2*kRegTap*16 bytes,
// 16-byte aligned) holds the per-sample tap tables meanwhile and is free again on return. Returns the RoI's
// feature-column range [cmin, cmin + ncols), ncols even (padded with a zero-weight column).
template <int PH, int PW, int NT>
__device__ __forceinline__ void reg_build_tables(RegTables<(PH + 1) / 2>& tb, void* scratch, const RoiGeom& g,
                                                 const int H, const int W, int& cmin, int& ncols, int& creal0,
                                                 int& creal1, int (&grp0)[2], int (&grpn)[2]) {
    constexpr int NU = (PH + 1) / 2;
    RXTap* xs = reinterpret_cast<RXTap*>(scratch);
    RXTap* ys = xs + kRegTap;
    const int tid = threadIdx.x, lane = tid & 31;
    const int gh = g.grid_h, gw = g.grid_w;
    const bool empty = gh <= 0 || gw <= 0;
    const bool tables = !empty && (long long)PW * gw <= kRegTap && (long long)PH * gh <= kRegTap;
    const float rcount = 1.0f / g.count;
    cmin = 0; ncols = 0; creal0 = 0; creal1 = 0;
    grp0[0] = grp0[1] = grpn[0] = grpn[1] = 0;
    if (tid == 0) tb.mode = empty ? 2 : (tables ? 0 : 1);
    if (tables) {
        const int nx = PW * gw, ny = PH * gh;
        for (int s = tid; s < nx; s += NT) {
            const int pw = s / gw;
            xs[s] = reg_xtap(g.start_w, g.bin_w, pw, s - pw * gw, gw, W);
        }
        for (int s = NT - 1 - tid; s < ny; s += NT) {
            const int ph = s / gh;
            ys[s] = reg_xtap(g.start_h, g.bin_h, ph, s - ph * gh, gh, H);
        }
    }
    __syncthreads();
    if (!tables) return;
    // merged y tables: thread (unit, entry e, row r) sums the weights of row 2u + r's samples on feature row ymin + e.
    // The rows a unit touches are consecutive (sample positions are monotone), so the table is just [ymin, ymin + n).
    for (int idx = tid; idx < NU * kRegYEnt * 2; idx += NT) {
        const int r = idx & 1, e = (idx >> 1) % kRegYEnt, u = idx / (2 * kRegYEnt);
        const int s0 = u * 2 * gh, s1 = min(PH, u * 2 + 2) * gh;
        int ylo = INT_MAX, yhi = -1;
        for (int sidx = s0; sidx < s1; ++sidx) {
            const RXTap Y = ys[sidx];
            if (Y.lo >= 0) { ylo = min(ylo, Y.lo); yhi = max(yhi, Y.hi); }
        }
        const int n = yhi < 0 ? 0 : yhi - ylo + 1;
        float w = 0.0f;
        if (e < n && u * 2 + r < PH) {
            const int y = ylo + e;
            for (int iy = 0; iy < gh; ++iy) {
                const RXTap Y = ys[(u * 2 + r) * gh + iy];
                if (Y.lo < 0) continue;
                if (Y.lo == y) w += Y.h;
                if (Y.hi == y) w += Y.l;
            }
        }
        if (r == 0) tb.yw[u][e].x = w; else tb.yw[u][e].y = w;
        if (e == 0 && r == 0) {
            tb.ymin[u] = yhi < 0 ? 0 : ylo;
            tb.ycnt[u] = n;
            if (n > kRegYEnt) tb.mode = 1;
        }
    }
    // feature-column range of the RoI (every warp computes it for itself)
    int lo = INT_MAX, hi = -1;
    for (int s = lane; s < PW * gw; s += 32) {
        const RXTap X = xs[s];
        if (X.lo >= 0) { lo = min(lo, X.lo); hi = max(hi, X.hi); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    const int rlo = lo, rhi = hi;                     // the columns that carry weight
    // the same for the two bin groups the forward walks separately (bins [0, kSplit) and [kSplit, PW))
    constexpr int kSplit = 2 * ((((PW + 1) / 2) + 1) / 2);
    int glo[2] = {INT_MAX, INT_MAX}, ghi[2] = {-1, -1};
    for (int s = lane; s < PW * gw; s += 32) {
        const RXTap X = xs[s];
        const int gi = s >= kSplit * gw;
        if (X.lo >= 0) { glo[gi] = min(glo[gi], X.lo); ghi[gi] = max(ghi[gi], X.hi); }
    }
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
        glo[gi] = __reduce_min_sync(0xffffffffu, glo[gi]);
        ghi[gi] = __reduce_max_sync(0xffffffffu, ghi[gi]);
    }
    if (hi >= 0 && ((hi - lo + 1) & 1)) {             // even column count: the forward walks two columns per step
        if (hi + 1 < W) ++hi; else if (lo > 0) --lo;  // (the added column has zero weights; W == 1: mode 1)
    }
    if (hi < 0) {
        if (tid == 0) tb.mode = 2;                    // every x sample lies outside the map
    } else if (hi - lo + 1 > kRegCols || ((hi - lo + 1) & 1)) {
        if (tid == 0) tb.mode = 1;
    } else {
        cmin = lo; ncols = hi - lo + 1;
        creal0 = rlo - lo; creal1 = rhi - lo + 1;
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
            if (ghi[gi] < 0) continue;                // no valid sample in this group of bins
            int a = glo[gi] - lo, b = ghi[gi] - lo;   // inclusive, relative to cmin; made even inside [0, ncols)
            if ((b - a + 1) & 1) { if (b + 1 < ncols) ++b; else --a; }
            grp0[gi] = a; grpn[gi] = b - a + 1;
        }
        for (int idx = tid; idx < ncols * 16; idx += NT) {   // dense x weights
            const int ci = idx >> 4, pw = idx & 15, col = cmin + ci;
            float w = 0.0f;
            if (pw < PW)
                for (int ix = 0; ix < gw; ++ix) {
                    const RXTap X = xs[pw * gw + ix];
                    if (X.lo < 0) continue;
                    if (X.lo == col) w += X.h;
                    if (X.hi == col) w += X.l;
                }
            tb.wxd[ci][pw] = w * rcount;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// x phase of one feature column for the NPG bin pairs starting at pair P0: acc[r][i] += Wx[col][2*(P0+i)..] * t[r]
template <int P0, int NPG, typename TB>
__device__ __forceinline__ void fwd_xphase(float2 (&acc)[2][NPG], const TB& tb, const int ci, const float2 t2) {
    const float* wrow = tb.wxd[ci] + 2 * P0;
#pragma unroll
    for (int i = 0; i < NPG; i += 2) {
        if (i + 1 < NPG) {
            const float4 x = *reinterpret_cast<const float4*>(wrow + 2 * i);
            acc[0][i] = ffma2(make_float2(x.x, x.y), t2.x, acc[0][i]);
            acc[1][i] = ffma2(make_float2(x.x, x.y), t2.y, acc[1][i]);
            acc[0][i + 1] = ffma2(make_float2(x.z, x.w), t2.x, acc[0][i + 1]);
            acc[1][i + 1] = ffma2(make_float2(x.z, x.w), t2.y, acc[1][i + 1]);
        } else {
            const float2 x = *reinterpret_cast<const float2*>(wrow + 2 * i);
            acc[0][i] = ffma2(x, t2.x, acc[0][i]);
            acc[1][i] = ffma2(x, t2.y, acc[1][i]);
        }
    }
}

// One pass over the feature columns [c0, c0 + ncols) (ncols even) at p0 for one unit, one group of bins and NJ
// consecutive feature rows (row stride rs): two columns per step, 2 * NJ independent loads in flight. CPL: channels per
// lane - with 2 a lane owns the adjacent channels 2*lane, 2*lane+1 of a 64-channel slab (64-bit loads) and everything
// that is not an FMA (loads, weight LDS, addresses, loop, per-slab setup) is shared by the two.
// NQ columns at p0: all NJ * NQ loads first, then the y and x phases
template <int CS, int NJ, int NQ, int P0, int NPG, int CPL, typename TB>
__device__ __forceinline__ void fwd_step(float2 (&acc)[CPL][2][NPG], const TB& tb, const float2 (&yw)[4],
                                         const float* __restrict__ p0, const int rs, const int ci, const int cstride) {
    float v[NQ][NJ][CPL];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (CPL == 2) {
                const float2 x = __ldg(reinterpret_cast<const float2*>(p0 + j * rs + q * cstride));
                v[q][j][0] = x.x; v[q][j][CPL - 1] = x.y;
            } else {
                v[q][j][0] = __ldg(p0 + j * rs + q * cstride);
            }
        }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int ch = 0; ch < CPL; ++ch) {
            float2 t2 = make_float2(yw[0].x * v[q][0][ch], yw[0].y * v[q][0][ch]);
#pragma unroll
            for (int j = 1; j < NJ; ++j) t2 = ffma2(yw[j], v[q][j][ch], t2);
            fwd_xphase<P0, NPG>(acc[ch], tb, ci + q, t2);
        }
}

template <int CS, int NJ, int P0, int NPG, int CPL, int MLP, typename TB>
__device__ __forceinline__ void fwd_columns(float2 (&acc)[CPL][2][NPG], const TB& tb, const float2 (&yw_r)[4],
                                            const float2* __restrict__ yw_s, const float* __restrict__ p0, const int rs,
                                            const int c0, const int ncols, const int C) {
    const int cstride = CS ? CS : C;
    // columns per step = independent 128-byte loads in flight per warp: 8..12 is the measured optimum (more spills,
    // less leaves the L2 latency exposed). MLP 0: NJ 1..3 -> 4 columns, NJ 4 -> 2; MLP 1: NJ 1 -> 8, NJ 2 -> 6.
    constexpr int NQ = MLP == 1 ? (NJ == 1 ? 8 : NJ == 2 ? 6 : NJ == 3 ? 4 : 2) : (NJ <= 3 ? 4 : 2);
    int ci = c0;
    const int cend = c0 + ncols;
#pragma unroll 1
    for (; ci + NQ <= cend; ci += NQ) {
        fwd_step<CS, NJ, NQ, P0, NPG, CPL>(acc, tb, yw_r, p0, rs, ci, cstride);
        p0 += NQ * cstride;
    }
    if (NQ > 4 && ci + 4 <= cend) {   // ncols is even: the rest is 0, 2, 4 (or 6 after an 8-step)
        fwd_step<CS, NJ, 4, P0, NPG, CPL>(acc, tb, yw_r, p0, rs, ci, cstride);
        p0 += 4 * cstride;
        ci += 4;
    }
    if (NQ > 2 && ci < cend) fwd_step<CS, NJ, 2, P0, NPG, CPL>(acc, tb, yw_r, p0, rs, ci, cstride);
}

// The same for units that touch 5..8 feature rows (RoIs taller than ~14 cells: two or three samples per bin along y):
// one column per step with all NJ rows' loads in flight, so the x phase - the bulk of the arithmetic - runs once per
// column instead of once per chunk of 4 rows. Rows 0..3 take their weights from registers, the rest from the table.
template <int CS, int NJ, int P0, int NPG, int CPL, int MLP, typename TB>
__device__ __forceinline__ void fwd_columns_tall(float2 (&acc)[CPL][2][NPG], const TB& tb, const float2 (&yw_r)[4],
                                                 const float2* __restrict__ yw_s, const float* __restrict__ p0, const int rs,
                                                 const int c0, const int ncols, const int C) {
    static_assert(NJ > 4 && NJ <= 8, "5..8 merged rows");
    const int cstride = CS ? CS : C;
    if (MLP == 1 && NJ <= 6) {   // two columns per step: 10..12 loads in flight
#pragma unroll 1
        for (int ci = c0; ci < c0 + ncols; ci += 2) {
            float v[2][NJ];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < NJ; ++j) v[q][j] = __ldg(p0 + j * rs + q * cstride);
            p0 += 2 * cstride;
            float2 yw[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) yw[j] = j < 4 ? yw_r[j & 3] : yw_s[j];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float2 t2 = make_float2(yw[0].x * v[q][0], yw[0].y * v[q][0]);
#pragma unroll
                for (int j = 1; j < NJ; ++j) t2 = ffma2(yw[j], v[q][j], t2);
                fwd_xphase<P0, NPG>(acc[0], tb, ci + q, t2);
            }
        }
        return;
    }
#pragma unroll 1
    for (int ci = c0; ci < c0 + ncols; ++ci) {
        float v[NJ][CPL];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (CPL == 2) {
                const float2 x = __ldg(reinterpret_cast<const float2*>(p0 + j * rs));
                v[j][0] = x.x; v[j][CPL - 1] = x.y;
            } else {
                v[j][0] = __ldg(p0 + j * rs);
            }
        }
        p0 += cstride;
        float2 yw[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) yw[j] = j < 4 ? yw_r[j & 3] : yw_s[j];
#pragma unroll
        for (int ch = 0; ch < CPL; ++ch) {
            float2 t2 = make_float2(yw[0].x * v[0][ch], yw[0].y * v[0][ch]);
#pragma unroll
            for (int j = 1; j < NJ; ++j) t2 = ffma2(yw[j], v[j][ch], t2);
            fwd_xphase<P0, NPG>(acc[ch], tb, ci, t2);
        }
    }
}

// all merged feature rows of the unit for one group of bins, <= 8 rows per pass over the columns (the first pass's
// first four weights are in registers across the CTA's slabs)
template <int CS, int P0, int NPG, int CPL, int MLP, typename TB>
__device__ __forceinline__ void fwd_group(float2 (&acc)[CPL][2][NPG], const TB& tb, const int u, const int ne,
                                          const float2 (&yw0)[4], const float* __restrict__ fcol, const int rs, const int c0,
                                          const int ncols, const int C) {
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < NPG; ++i) acc[ch][r][i] = make_float2(0.0f, 0.0f);
    if (ne <= 0 || ncols <= 0) return;
    const int cstride = CS ? CS : C;
    const float* __restrict__ p0 = fcol + (size_t)c0 * cstride;
    for (int e0 = 0; e0 < ne; e0 += 8) {   // one pass for all but whole-map RoIs
        const int n = min(ne - e0, 8);
        float2 yw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) yw[j] = e0 == 0 ? yw0[j] : (e0 + j < ne ? tb.yw[u][e0 + j] : make_float2(0.0f, 0.0f));
        const float2* __restrict__ yws = tb.yw[u] + e0;
        const float* __restrict__ pe = p0 + (size_t)e0 * rs;
        switch (n) {   // warp-uniform
            case 1: fwd_columns<CS, 1, P0, NPG, CPL, MLP>(acc, tb, yw, yws, pe, rs, c0, ncols, C); break;
            case 2: fwd_columns<CS, 2, P0, NPG, CPL, MLP>(acc, tb, yw, yws, pe, rs, c0, ncols, C); break;
            case 3: fwd_columns<CS, 3, P0, NPG, CPL, MLP>(acc, tb, yw, yws, pe, rs, c0, ncols, C); break;
            case 4: fwd_columns<CS, 4, P0, NPG, CPL, MLP>(acc, tb, yw, yws, pe, rs, c0, ncols, C); break;
            case 5: fwd_columns_tall<CS, 5, P0, NPG, CPL, (CPL == 1 ? MLP : 0)>(acc, tb, yw, yws, pe, rs, c0, ncols, C); break;
            case 6: fwd_columns_tall<CS, 6, P0, NPG, CPL, (CPL == 1 ? MLP : 0)>(acc, tb, yw, yws, pe, rs, c0, ncols, C); break;
            case 7: fwd_columns_tall<CS, 7, P0, NPG, CPL, (CPL == 1 ? MLP : 0)>(acc, tb, yw, yws, pe, rs, c0, ncols, C); break;
            default: fwd_columns_tall<CS, 8, P0, NPG, CPL, (CPL == 1 ? MLP : 0)>(acc, tb, yw, yws, pe, rs, c0, ncols, C); break;
        }
    }
}

// the group's accumulators -> the unit's rows of the [channels][PH*PW] tile (tp points at row 2u, bin 0 of the channel)
template <int PW, int P0, int NPG>
__device__ __forceinline__ void fwd_store_group(__half* __restrict__ tp, const float2 (&acc)[2][NPG], const int nr) {
    // fp16 output (the reference under autocast): fp32 accumulate, round to nearest even like the torchvision autocast path
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < NPG; ++i) {
            const int b = 2 * (P0 + i);
            if (PW % 2 == 0) {   // rows start 4-byte aligned: one half2 store per bin pair
                if (r < nr) *reinterpret_cast<__half2*>(tp + r * PW + b) = __floats2half2_rn(acc[r][i].x, acc[r][i].y);
            } else {
                if (r < nr && b < PW) tp[r * PW + b] = __float2half_rn(acc[r][i].x);
                if (r < nr && b + 1 < PW) tp[r * PW + b + 1] = __float2half_rn(acc[r][i].y);
            }
        }
}

template <int PW, int P0, int NPG>
__device__ __forceinline__ void fwd_store_group(float* __restrict__ tp, const float2 (&acc)[2][NPG], const int nr) {
    if constexpr (PW % 2 == 0 && (2 * PW) % 4 == 0 && P0 % 2 == 0) {
        // row 2u starts 16-byte aligned, row 2u+1 (PW floats later, PW = 2 mod 4) 8-byte aligned
#pragma unroll
        for (int i = 0; i < NPG; i += 2) {
            if (i + 1 < NPG)
                *reinterpret_cast<float4*>(tp + 2 * (P0 + i)) = make_float4(acc[0][i].x, acc[0][i].y, acc[0][i + 1].x, acc[0][i + 1].y);
            else
                *reinterpret_cast<float2*>(tp + 2 * (P0 + i)) = acc[0][i];
        }
#pragma unroll
        for (int i = 0; i < NPG; ++i) *reinterpret_cast<float2*>(tp + PW + 2 * (P0 + i)) = acc[1][i];
    } else {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < NPG; ++i) {
                const int b = 2 * (P0 + i);
                if (r < nr && b < PW) tp[r * PW + b] = acc[r][i].x;
                if (r < nr && b + 1 < PW) tp[r * PW + b + 1] = acc[r][i].y;
            }
    }
}

template <typename T, int PH, int PW, int CS, int OCC, int CPL, int MLP>
__global__ void __launch_bounds__(32 * ((PH + 1) / 2), OCC)
roi_align_fwd_reg_kernel(const RoiParams p, T* __restrict__ out, const int cgroups, const int slabs) {
    constexpr int NU = (PH + 1) / 2, NB = PH * PW, NT = 32 * NU, NP = (PW + 1) / 2, CC = 32 * CPL;
    constexpr int NPA = PW >= 12 ? (NP + 1) / 2 : NP, NPB = NP - NPA;   // bin pairs of the two groups (one for 7x7)
    extern __shared__ __align__(128) float tile_raw[];   // [CC channels][NB] of T: the CTA's contiguous output region
    T* tile = reinterpret_cast<T*>(tile_raw);            // (>= 4 KB: it doubles as tap-table scratch while the tables are built)
    __shared__ RegTables<NU> tb;

    const int kk = blockIdx.x / cgroups;
    if (p.k_dev && kk >= __ldg(p.k_dev)) return;   // capacity launch: RoI beyond the live count
    const int k = p.perm ? __ldg(p.perm + kk) : kk;   // launch order (coin_roi_launch_order): small RoIs last
    const int cg0 = (blockIdx.x - kk * cgroups) * (CC * slabs);
    const int lane = threadIdx.x & 31, u = threadIdx.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W;
    const int C = CS ? CS : p.C;
    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    const float* __restrict__ fimg = L.feat_nhwc + (size_t)g.batch * H * W * C;
    const int nslab = min(slabs, (C - cg0) / CC);
    T* __restrict__ oroi = out + (size_t)k * C * NB;

    int cmin, ncols, creal0, creal1, grp0[2], grpn[2];
    reg_build_tables<PH, PW, NT>(tb, tile_raw, g, H, W, cmin, ncols, creal0, creal1, grp0, grpn);
    const int mode = tb.mode;

    if (mode == 2) {   // no sample inside the map: the RoI pools to zeros
        T* o = oroi + (size_t)cg0 * NB;
        for (int e = threadIdx.x; e < nslab * CC * NB; e += NT) o[e] = from_f32<T>(0.0f);
        return;
    }
    if (mode == 1) {   // exotic geometry (sampling grids beyond the tables): direct 4-tap evaluation
        const float rcount = 1.0f / g.count;
        for (int sl = 0; sl < nslab * CPL; ++sl) {
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
                oroi[(size_t)c * NB + b] = from_f32<T>(acc * rcount);
            }
        }
        return;
    }

    // The unit's bins are walked in two groups (left and right half of the output row): a column of a wide RoI
    // carries weight on a few adjacent bins only, so each group needs its own, nearly disjoint, column range and
    // half the x-phase arithmetic and accumulator registers of a pass over all bins; narrow RoIs have few columns.
    const int ne = tb.ycnt[u], rs = W * C;
    float2 yw0[4];   // the unit's merged y table stays in registers across the CTA's channel slabs (first <= 4 rows)
#pragma unroll
    for (int j = 0; j < 4; ++j) yw0[j] = j < ne ? tb.yw[u][j] : make_float2(0.0f, 0.0f);
    const float* __restrict__ funit = fimg + (size_t)tb.ymin[u] * rs + (size_t)cmin * C + cg0 + CPL * lane;
    const int nr = min(2, PH - u * 2);
    const uint64_t pol = l2_evict_first_policy();
    T* __restrict__ tp = tile + CPL * lane * NB + u * (2 * PW);

    for (int sl = 0; sl < nslab; ++sl) {
        const float* __restrict__ fcol = funit + sl * CC;
        {
            float2 acc[CPL][2][NPA];
            fwd_group<CS, 0, NPA, CPL, MLP>(acc, tb, u, ne, yw0, fcol, rs, NPB > 0 ? grp0[0] : 0, NPB > 0 ? grpn[0] : ncols, C);
            // the previous slab's bulk store must have read the tile before it is overwritten
            if (sl > 0 && threadIdx.x == 0) reg_bulk_wait_read();
            __syncthreads();
#pragma unroll
            for (int ch = 0; ch < CPL; ++ch) fwd_store_group<PW, 0, NPA>(tp + ch * NB, acc[ch], nr);
        }
        if constexpr (NPB > 0) {
            float2 acc[CPL][2][NPB > 0 ? NPB : 1];
            fwd_group<CS, NPA, (NPB > 0 ? NPB : 1), CPL, MLP>(acc, tb, u, ne, yw0, fcol, rs, grp0[1], grpn[1], C);
#pragma unroll
            for (int ch = 0; ch < CPL; ++ch) fwd_store_group<PW, NPA, (NPB > 0 ? NPB : 1)>(tp + ch * NB, acc[ch], nr);
        }
        reg_fence_async();
        __syncthreads();
        if (threadIdx.x == 0) reg_bulk_store(oroi + (size_t)(cg0 + sl * CC) * NB, tile, CC * NB * sizeof(T), pol);
    }
    if (threadIdx.x == 0) reg_bulk_wait_read();
}

template <typename T, int PH, int PW, int CS, int OCC, int CPL, int MLP = 1>
static int launch_fwd_reg(const RoiParams& p, T* out, int slabs, cudaStream_t s) {
    constexpr int NU = (PH + 1) / 2, NB = PH * PW, CC = 32 * CPL;
    auto kern = roi_align_fwd_reg_kernel<T, PH, PW, CS, OCC, CPL, MLP>;
    const size_t smem = reg_occupancy_pad(std::max<size_t>((size_t)CC * NB * sizeof(T), 2 * kRegTap * sizeof(RXTap)));
    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    reg_apply_carveout(kern);
    slabs = std::max(1, std::min(slabs, p.C / CC));
    const int cgroups = (int)ceil_div(p.C, CC * slabs);
    kern<<<(unsigned)(p.K * cgroups), 32 * NU, smem, s>>>(p, out, cgroups, slabs);
    return check_launch("roi_align_fwd_reg_kernel");
}

// ------------------------------------------------------------------------------------------------
// backward: grad_in[y][x][c] += sum_{ph,pw} Wy[ph][y] * Wx[pw][x] / count * g[c][ph][pw]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void reg_bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(reg_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            reg_smem_u32(sdst)),
        "l"(gsrc), "r"(bytes), "r"(reg_smem_u32(bar)), "l"(pol)
        : "memory");
}

// x phase of the backward for one feature column: u_r = sum_pw Wx[col][pw] * g[r][pw]
template <int PW, typename TB>
__device__ __forceinline__ void bwd_xphase(const float2 (&gr)[2][(PW + 1) / 2], const TB& tb, const int ci, float& u0, float& u1) {
    constexpr int NP = (PW + 1) / 2;
    const float4* wrow = reinterpret_cast<const float4*>(tb.wxd[ci]);
    float2 s0 = make_float2(0.0f, 0.0f), s1 = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < (NP + 1) / 2; ++i) {
        const float4 x = wrow[i];
        s0 = ffma2v(gr[0][2 * i], make_float2(x.x, x.y), s0);
        s1 = ffma2v(gr[1][2 * i], make_float2(x.x, x.y), s1);
        if (2 * i + 1 < NP) {
            s0 = ffma2v(gr[0][2 * i + 1], make_float2(x.z, x.w), s0);
            s1 = ffma2v(gr[1][2 * i + 1], make_float2(x.z, x.w), s1);
        }
    }
    u0 = s0.x + s0.y;
    u1 = s1.x + s1.y;
}

// One pass over the RoI's feature columns [c0, c1) for one unit: per column the x phase, then one fp32 RED per merged
// feature row. The first NJ (<= 4) rows' weights are in registers; TALL units read the rest from the table. Two
// columns per iteration share the RED address arithmetic (the second column is an immediate offset).
template <int PW, int CS, int NJ, bool TALL, typename TB>
__device__ __forceinline__ void bwd_columns(const float2 (&gr)[2][(PW + 1) / 2], const TB& tb, const int u, const int ne,
                                            const float2 (&yw)[4], float* __restrict__ p0, const int rs, const int c0,
                                            const int c1, const int C) {
    const int cstride = CS ? CS : C;
    p0 += (size_t)c0 * cstride;
    int ci = c0;
#pragma unroll 1
    for (; ci + 2 <= c1; ci += 2) {
        float a0, a1, b0, b1;
        bwd_xphase<PW>(gr, tb, ci, a0, a1);
        bwd_xphase<PW>(gr, tb, ci + 1, b0, b1);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float* q = p0 + j * rs;
            atomicAdd(q, __fmaf_rn(yw[j].y, a1, yw[j].x * a0));
            atomicAdd(q + cstride, __fmaf_rn(yw[j].y, b1, yw[j].x * b0));
        }
        if (TALL)
            for (int j = 4; j < ne; ++j) {   // warp-uniform, rare
                const float2 w = tb.yw[u][j];
                float* q = p0 + (size_t)j * rs;
                atomicAdd(q, __fmaf_rn(w.y, a1, w.x * a0));
                atomicAdd(q + cstride, __fmaf_rn(w.y, b1, w.x * b0));
            }
        p0 += 2 * cstride;
    }
    if (ci < c1) {
        float a0, a1;
        bwd_xphase<PW>(gr, tb, ci, a0, a1);
#pragma unroll
        for (int j = 0; j < NJ; ++j) atomicAdd(p0 + j * rs, __fmaf_rn(yw[j].y, a1, yw[j].x * a0));
        if (TALL)
            for (int j = 4; j < ne; ++j) {
                const float2 w = tb.yw[u][j];
                atomicAdd(p0 + (size_t)j * rs, __fmaf_rn(w.y, a1, w.x * a0));
            }
    }
}

template <typename T, int PH, int PW, int CS, int OCC>
__global__ void __launch_bounds__(32 * ((PH + 1) / 2), OCC)
roi_align_bwd_reg_kernel(const RoiParams p, const T* __restrict__ go, const int cgroups, const int slabs) {
    constexpr int NU = (PH + 1) / 2, NB = PH * PW, NT = 32 * NU, NP = (PW + 1) / 2;
    constexpr bool VEC = PW % 2 == 0 && (2 * PW) % 4 == 0 && NB % 4 == 0;
    constexpr uint32_t kTileBytes = 32 * NB * sizeof(T);
    extern __shared__ __align__(128) float tile_raw[];   // [32 channels][NB] grad_out tile of T, then 4 KB of tap-table scratch
    const T* tile = reinterpret_cast<const T*>(tile_raw);
    __shared__ RegTables<NU> tb;
    __shared__ __align__(8) uint64_t bar_full;

    const int kk = blockIdx.x / cgroups;
    if (p.k_dev && kk >= __ldg(p.k_dev)) return;   // capacity launch: RoI beyond the live count
    const int k = p.perm ? __ldg(p.perm + kk) : kk;   // launch order (coin_roi_launch_order): small RoIs last
    const int cg0 = (blockIdx.x - kk * cgroups) * (32 * slabs);
    const int lane = threadIdx.x & 31, u = threadIdx.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W;
    const int C = CS ? CS : p.C;
    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    if (g.grid_h <= 0 || g.grid_w <= 0) return;    // no samples: no gradient
    float* __restrict__ gimg = const_cast<float*>(L.feat_nhwc) + (size_t)g.batch * H * W * C;
    const int nslab = min(slabs, (C - cg0) / 32);
    const T* __restrict__ groi = go + ((size_t)k * C + cg0) * NB;
    const uint64_t pol = l2_evict_first_policy();

    if (threadIdx.x == 0) {   // the first tile streams in while the tables are built
        reg_mbar_init(&bar_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        reg_bulk_load(tile_raw, groi, kTileBytes, &bar_full, pol);
    }
    int cmin, ncols, creal0, creal1, grp0[2], grpn[2];
    reg_build_tables<PH, PW, NT>(tb, reinterpret_cast<char*>(tile_raw) + kTileBytes, g, H, W, cmin, ncols, creal0, creal1, grp0, grpn);
    const int mode = tb.mode;

    if (mode == 2) {   // no sample inside the map: no gradient (the tile in flight must land before the CTA may exit)
        reg_mbar_wait(&bar_full, 0);
        return;
    }
    if (mode == 1) {   // exotic geometry: direct 4-tap scatter
        reg_mbar_wait(&bar_full, 0);
        const float rcount = 1.0f / g.count;
        for (int sl = 0; sl < nslab; ++sl) {
            const int c = cg0 + sl * 32 + lane;
            for (int b = u; b < NB; b += NU) {
                const int ph = b / PW, pw = b - ph * PW;
                const float gv = to_f32(__ldg(groi + (size_t)(sl * 32 + lane) * NB + b)) * rcount;
                for (int iy = 0; iy < g.grid_h; ++iy) {
                    const Tap Y = make_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H, W * C);
                    if (Y.lo < 0) continue;
                    for (int ix = 0; ix < g.grid_w; ++ix) {
                        const Tap X = make_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W, C);
                        if (X.lo < 0) continue;
                        float* f = gimg + c;
                        atomicAdd(f + Y.lo + X.lo, Y.h * X.h * gv);
                        atomicAdd(f + Y.lo + X.hi, Y.h * X.l * gv);
                        atomicAdd(f + Y.hi + X.lo, Y.l * X.h * gv);
                        atomicAdd(f + Y.hi + X.hi, Y.l * X.l * gv);
                    }
                }
            }
        }
        return;
    }

    const int ne = tb.ycnt[u], rs = W * C;
    float2 yw0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) yw0[j] = j < ne ? tb.yw[u][j] : make_float2(0.0f, 0.0f);
    float* __restrict__ gunit = gimg + (size_t)tb.ymin[u] * rs + (size_t)cmin * C + cg0 + lane;
    const int nr = min(2, PH - u * 2);
    const T* __restrict__ tp = tile + lane * NB + u * (2 * PW);

    for (int sl = 0; sl < nslab; ++sl) {
        reg_mbar_wait(&bar_full, sl & 1);          // the slab's [32][NB] grad_out tile has landed
        float2 gr[2][NP];
        if constexpr (sizeof(T) == 2) {   // fp16 gradients (the reference under autocast): widened to fp32 here
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const int b0 = 2 * i, b1 = 2 * i + 1;
                    if (PW % 2 == 0) {
                        gr[r][i] = r < nr ? __half22float2(*reinterpret_cast<const __half2*>(tp + r * PW + b0)) : make_float2(0.0f, 0.0f);
                    } else {
                        gr[r][i].x = (r < nr && b0 < PW) ? to_f32(tp[r * PW + b0]) : 0.0f;
                        gr[r][i].y = (r < nr && b1 < PW) ? to_f32(tp[r * PW + b1]) : 0.0f;
                    }
                }
        } else if (VEC) {   // rows 2u, 2u+1 are 2*PW consecutive floats of the channel's plane
            const float4* tp4 = reinterpret_cast<const float4*>(tp);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const float4 x = tp4[i];
                const int r = (2 * i) / NP, a = 2 * i - r * NP;
                const int r2 = (2 * i + 1) / NP, b = 2 * i + 1 - r2 * NP;
                gr[r][a] = make_float2(x.x, x.y);
                gr[r2][b] = make_float2(x.z, x.w);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const int b0 = 2 * i, b1 = 2 * i + 1;
                    gr[r][i].x = (r < nr && b0 < PW) ? to_f32(tp[r * PW + b0]) : 0.0f;
                    gr[r][i].y = (r < nr && b1 < PW) ? to_f32(tp[r * PW + b1]) : 0.0f;
                }
        }
        __syncthreads();                            // every warp holds its rows in registers: the tile is free
        if (threadIdx.x == 0 && sl + 1 < nslab)     // the next tile streams in under this slab's arithmetic
            reg_bulk_load(tile_raw, groi + (size_t)(sl + 1) * 32 * NB, kTileBytes, &bar_full, pol);
        if (ne > 0) {
            float* __restrict__ gcol = gunit + sl * 32;
            switch (min(ne, 5)) {   // warp-uniform
                case 1: bwd_columns<PW, CS, 1, false>(gr, tb, u, ne, yw0, gcol, rs, creal0, creal1, C); break;
                case 2: bwd_columns<PW, CS, 2, false>(gr, tb, u, ne, yw0, gcol, rs, creal0, creal1, C); break;
                case 3: bwd_columns<PW, CS, 3, false>(gr, tb, u, ne, yw0, gcol, rs, creal0, creal1, C); break;
                case 4: bwd_columns<PW, CS, 4, false>(gr, tb, u, ne, yw0, gcol, rs, creal0, creal1, C); break;
                default: bwd_columns<PW, CS, 4, true>(gr, tb, u, ne, yw0, gcol, rs, creal0, creal1, C); break;
            }
        }
    }
}

template <typename T, int PH, int PW, int CS, int OCC>
static int launch_bwd_reg(const RoiParams& p, const T* go, int slabs, cudaStream_t s) {
    constexpr int NU = (PH + 1) / 2, NB = PH * PW;
    auto kern = roi_align_bwd_reg_kernel<T, PH, PW, CS, OCC>;
    const size_t smem = reg_occupancy_pad((size_t)32 * NB * sizeof(T) + 2 * kRegTap * sizeof(RXTap));
    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    reg_apply_carveout(kern);
    const int cgroups = (int)ceil_div(p.C, 32 * slabs);
    kern<<<(unsigned)(p.K * cgroups), 32 * NU, smem, s>>>(p, go, cgroups, slabs);
    return check_launch("roi_align_bwd_reg_kernel");
}

bool roi_align_bwd_reg_supported(const RoiParams& p, int grad_dtype) {
    if (reg_env("COIN_ROI_REG", 1) == 0 || reg_env("COIN_ROI_BWD_REG", 1) == 0) return false;
    if ((grad_dtype != COIN_F32 && grad_dtype != COIN_F16) || p.C % 32 != 0) return false;
    return (p.PH == 14 && p.PW == 14) || (p.PH == 7 && p.PW == 7);
}

template <typename T>
static int dispatch_bwd_reg(const RoiParams& p, const T* g, cudaStream_t s) {
    const int nsl = (int)(p.C / 32);
    int slabs = reg_env("COIN_ROI_BWD_REG_SLABS", p.K < 1024 ? 2 : (p.PH == 7 ? 4 : 8));
    slabs = std::max(1, std::min(slabs, nsl));
    if (p.PH == 14) {
        if (p.C == 1024) return launch_bwd_reg<T, 14, 14, 1024, 4>(p, g, slabs, s);
        return launch_bwd_reg<T, 14, 14, 0, 4>(p, g, slabs, s);
    }
    if (p.C == 1024) return launch_bwd_reg<T, 7, 7, 1024, 8>(p, g, slabs, s);
    return launch_bwd_reg<T, 7, 7, 0, 8>(p, g, slabs, s);
}

int launch_roi_align_bwd_reg(const RoiParams& p, const void* grad_out, int grad_dtype, cudaStream_t s) {
    if (reinterpret_cast<uintptr_t>(grad_out) & 15) return fail(COIN_ERR_INVALID, "roi_align_bwd: grad_out must be 16-byte aligned");
    if (grad_dtype == COIN_F16) return dispatch_bwd_reg<__half>(p, static_cast<const __half*>(grad_out), s);
    return dispatch_bwd_reg<float>(p, static_cast<const float*>(grad_out), s);
}

bool roi_align_fwd_reg_supported(const RoiParams& p, int out_dtype) {
    if (reg_env("COIN_ROI_REG", 1) == 0) return false;
    if ((out_dtype != COIN_F32 && out_dtype != COIN_F16) || p.C % 32 != 0) return false;
    // COIN_ROI_REG_MINK = n: capacity launches (device-side RoI count) of fewer than n RoIs take the separable kernel. Off by
    // default: what made such launches slow were map-sized RoIs (one CTA walking 256 channels of the whole map), and callers
    // now divert those to the separable kernel themselves (ops.roi_align_forward_planned, coin_roi_split_by_area).
    if (p.K < reg_env("COIN_ROI_REG_MINK", 0) && p.k_dev) return false;
    return (p.PH == 14 && p.PW == 14) || (p.PH == 7 && p.PW == 7);
}

template <typename T>
static int dispatch_fwd_reg(const RoiParams& p, T* o, cudaStream_t s) {
    // two channels per lane (64-channel slabs, 64-bit loads): measured slower than one (3 CTAs/SM), kept selectable
    const int cpl = (p.C % 64 == 0 && (reinterpret_cast<uintptr_t>(p.lv[0].feat_nhwc) & 7) == 0) ? reg_env("COIN_ROI_REG_CPL", 1) : 1;
    // channels per CTA: 256 (the tables are built once per CTA; 128 for the cheaper 7x7 units); few RoIs: fewer, so
    // that one very large RoI cannot leave a long tail (measured: foggy 14x14 256 -> 372 us, 128 -> 380, 512 -> 399)
    const int chans = p.K < 1024 ? reg_env("COIN_ROI_REG_CHANS_SMALL", 64) : reg_env("COIN_ROI_REG_CHANS", p.PH == 7 ? 128 : 256);
    if (p.PH == 14) {
        if (p.C == 1024) {
            if (cpl == 2) return launch_fwd_reg<T, 14, 14, 1024, 3, 2>(p, o, chans / 64, s);
            if (reg_env("COIN_ROI_REG_MLP", 1) == 0) return launch_fwd_reg<T, 14, 14, 1024, 4, 1, 0>(p, o, chans / 32, s);
            return launch_fwd_reg<T, 14, 14, 1024, 4, 1>(p, o, chans / 32, s);
        }
        if (cpl == 2) return launch_fwd_reg<T, 14, 14, 0, 3, 2>(p, o, chans / 64, s);
        return launch_fwd_reg<T, 14, 14, 0, 4, 1>(p, o, chans / 32, s);
    }
    if (p.C == 1024) {
        if (cpl == 2) return launch_fwd_reg<T, 7, 7, 1024, 6, 2>(p, o, chans / 64, s);
        return launch_fwd_reg<T, 7, 7, 1024, 8, 1>(p, o, chans / 32, s);
    }
    if (cpl == 2) return launch_fwd_reg<T, 7, 7, 0, 6, 2>(p, o, chans / 64, s);
    return launch_fwd_reg<T, 7, 7, 0, 8, 1>(p, o, chans / 32, s);
}

int launch_roi_align_fwd_reg(const RoiParams& p, void* out, int out_dtype, cudaStream_t s) {
    if (reinterpret_cast<uintptr_t>(out) & 15) return fail(COIN_ERR_INVALID, "roi_align_fwd: out must be 16-byte aligned");
    if (out_dtype == COIN_F16) return dispatch_fwd_reg<__half>(p, static_cast<__half*>(out), s);
    return dispatch_fwd_reg<float>(p, static_cast<float*>(out), s);
}

}  // namespace coin
