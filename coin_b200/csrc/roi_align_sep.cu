// roi_align_sep.cu -- separable ROIAlign forward for sm_100a (the default, non-parity-mode kernel).
//
// Replaces torchvision::roi_align as reached from coin/modeling/roi_heads/clip_roi_heads.py:51-63,
// 142-147,172-176 (ROIPooler -> ROIAlign). Same sample positions, bilinear weights, validity rule and
// 1/count scaling as the torchvision kernel; the summation is re-associated (see below), so results
// agree to ~1e-7 relative instead of bit for bit. COIN_ROI_EXACT=1 selects the bit-exact kernel of
// roi_align.cu instead.
//
// ROIAlign is a separable linear map per RoI and channel:
//     out[ph,pw] = 1/count * sum_{iy,ix} bilinear(F, y(ph,iy), x(pw,ix))
//                = 1/count * sum_ix ( hx * T[ph][xlo] + lx * T[ph][xhi] ),
//     T[ph][x]   = sum_iy ( hy * F[ylo][x] + ly * F[yhi][x] )            (valid samples only;
// a sample is valid iff its y is valid and its x is valid, so validity factorises too).
//
// One warp owns a "unit" = a group of output rows x a chunk of output columns x a 32*CPL-channel slab:
//   phase 1 (lane = channel): T for the unit's rows and feature columns, from fp32 NHWC global memory.
//           Every load is a coalesced 128-byte line; the iterations over feature columns are
//           independent, so a warp keeps 2*U*CPL loads in flight (the v2 kernel stalled on one
//           dependent L2 round trip per sample). T goes to a per-warp shared-memory buffer with an odd
//           pitch.
//   phase 2 (lane = output bin): each lane holds the <= 4 column weights of its bin in registers and
//           walks the channels: <= 4 conflict-free LDS + FMA per output, and the warp's store covers
//           nrows*PW consecutive floats of out[k, c, :, :] -- coalesced without staging the output
//           tile in shared memory.
// Both phases are warp-local (no CTA barrier after the tap tables are built). A CTA loops over
// several channel slabs of the same RoI so the tables are built once per 256 channels.
#include <climits>

#include "roi_common.cuh"

namespace coin {

static inline int sep_env(const char* name, int dflt) { return option(name, dflt); }

constexpr int kSepTap = 128;     // tap-table entries per axis (PW*grid_w and PH*grid_h must fit)
constexpr int kSepCells = 32;    // T cells (row x feature column) per unit
constexpr int kSepChunks = 16;   // max column chunks per RoI

struct SepChunk { int pa, nb, ca, ncol; };  // bins [pa, pa+nb) need feature columns [ca, ca+ncol)

struct XTap { int lo, hi; float l, h; };    // cell indices (not offsets); lo < 0: sample outside the map

__device__ __forceinline__ XTap make_xtap(float start, float bin, int p, int i, int grid, int size) {
    const float v = start + (float)p * bin + ((float)i + 0.5f) * bin / (float)grid;
    XTap t;
    if (!axis_taps(v, size, t.lo, t.hi, t.l, t.h)) { t.lo = -1; t.hi = -1; t.l = 0.0f; t.h = 0.0f; }
    return t;
}

// Phase 1 for one output row: T[col][channel] = sum_iy hy*F[ylo][col] + ly*F[yhi][col] for the columns
// [0, ncol) of the chunk (fb already points at the chunk's first column and this lane's channel).
// U columns are processed per step with all 2*U*CPL loads issued before the first use. FAST: full channel
// slab and ncol >= U: the last step is shifted back to end at ncol (recomputing a few columns) so no
// load needs a predicate.
template <int U, int CPL, bool FAST>
__device__ __forceinline__ void sep_phase1_row(const float* __restrict__ fb, const Tap* __restrict__ yt, const int gh,
                                               const int ncol, const int C, const int cc, const int lane,
                                               float* __restrict__ trow) {
    constexpr int P = 32 * CPL + 1;
    for (int col0 = 0; col0 < ncol; col0 += U) {
        const int cs = FAST ? min(col0, ncol - U) : col0;
        float t[U][CPL];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < CPL; ++j) t[u][j] = 0.0f;
        for (int iy = 0; iy < gh; ++iy) {
            const Tap Y = yt[iy];
            if (Y.lo < 0) continue;                       // warp-uniform
            const float* __restrict__ rl = fb + Y.lo + (size_t)cs * C;
            const float* __restrict__ rh = fb + Y.hi + (size_t)cs * C;
            float a[U][CPL], b[U][CPL];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    if (FAST) {
                        a[u][j] = __ldg(rl + u * C + 32 * j);
                        b[u][j] = __ldg(rh + u * C + 32 * j);
                    } else {
                        const bool ok = cs + u < ncol && lane + 32 * j < cc;
                        a[u][j] = ok ? __ldg(rl + u * C + 32 * j) : 0.0f;
                        b[u][j] = ok ? __ldg(rh + u * C + 32 * j) : 0.0f;
                    }
                }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < CPL; ++j) t[u][j] = __fmaf_rn(Y.l, b[u][j], __fmaf_rn(Y.h, a[u][j], t[u][j]));
        }
        float* __restrict__ tw = trow + cs * P;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (FAST || cs + u < ncol) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) tw[u * P + 32 * j] = t[u][j];
            }
    }
}

// Phase 2 inner loop: NT column taps per output, `n` channels starting at the lane's base pointers.
template <typename OutT, int NT, int P>
__device__ __forceinline__ void sep_phase2_loop(const float* __restrict__ tb, const int o1, const int o2, const int o3,
                                                const float w0, const float w1, const float w2, const float w3,
                                                OutT* __restrict__ ob, const int NB, const int n) {
    constexpr int UN = NT <= 2 ? 8 : 4;
    int c = 0;
    for (; c + UN <= n; c += UN) {
#pragma unroll
        for (int q = 0; q < UN; ++q) {
            float acc = w0 * tb[c + q];
            acc = __fmaf_rn(w1, tb[o1 + c + q], acc);
            if (NT >= 3) acc = __fmaf_rn(w2, tb[o2 + c + q], acc);
            if (NT >= 4) acc = __fmaf_rn(w3, tb[o3 + c + q], acc);
            ob[(size_t)(c + q) * NB] = from_f32<OutT>(acc);
        }
    }
    for (; c < n; ++c) {
        float acc = w0 * tb[c];
        acc = __fmaf_rn(w1, tb[o1 + c], acc);
        if (NT >= 3) acc = __fmaf_rn(w2, tb[o2 + c], acc);
        if (NT >= 4) acc = __fmaf_rn(w3, tb[o3 + c], acc);
        ob[(size_t)c * NB] = from_f32<OutT>(acc);
    }
}

// CS: channel stride of the NHWC map in elements when known at compile time (0 = use p.C); PHT/PWT: the
// output size when known at compile time (0 = use p.PH / p.PW). Compile-time strides turn the address
// arithmetic of the two inner loops into immediate offsets.
template <typename OutT, int CPL, int CS, int PHT, int PWT>
__global__ void __launch_bounds__(256, 3)
roi_align_fwd_sep_kernel(const RoiParams p, OutT* __restrict__ out, const int cgroups, const int slabs) {
    constexpr int CC = 32 * CPL, P = CC + 1;
    extern __shared__ float tsm[];                 // [nwarps][kSepCells][P]
    __shared__ XTap xs[kSepTap];
    __shared__ Tap ys[kSepTap];
    __shared__ int blo[kSepTap], bhi[kSepTap];
    __shared__ SepChunk chunks[kSepChunks];
    __shared__ int s_nchunks, s_rows, s_mode;      // mode 0: separable, 1: direct, 2: all zero

    const int kk = blockIdx.x / cgroups;
    if (p.k_dev && kk >= __ldg(p.k_dev)) return;   // capacity launch: RoI beyond the live count
    const int k = p.perm ? __ldg(p.perm + kk) : kk;   // launch order / RoI subset (coin_roi_launch_order, coin_roi_split_by_area)
    const int cg0 = (blockIdx.x - kk * cgroups) * (CC * slabs);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W;
    const int C = CS ? CS : p.C, PH = PHT ? PHT : p.PH, PW = PWT ? PWT : p.PW, NB = PH * PW;
    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    const float* __restrict__ fimg = L.feat_nhwc + (size_t)g.batch * H * W * C;
    const float rcount = 1.0f / g.count;
    const int nslab = min(slabs, (int)((C - cg0 + CC - 1) / CC));
    OutT* __restrict__ oroi = out + (size_t)k * C * NB;

    const bool empty = g.grid_h <= 0 || g.grid_w <= 0;
    const bool tables = !empty && PW <= kSepTap && PH <= kSepTap && (long long)PW * g.grid_w <= kSepTap &&
                        (long long)PH * g.grid_h <= kSepTap;
    if (tables) {   // x taps from the low thread ids, y taps from the high ones: both tables fill concurrently
        const int nx = PW * g.grid_w, ny = PH * g.grid_h;
        for (int s = threadIdx.x; s < nx; s += blockDim.x) {
            const int pw = s / g.grid_w;
            xs[s] = make_xtap(g.start_w, g.bin_w, pw, s - pw * g.grid_w, g.grid_w, W);
        }
        for (int s = blockDim.x - 1 - threadIdx.x; s < ny; s += blockDim.x) {
            const int ph = s / g.grid_h;
            ys[s] = make_tap(g.start_h, g.bin_h, ph, s - ph * g.grid_h, g.grid_h, H, W * C);
        }
    }
    __syncthreads();
    if (warp == 0) {   // chunk list: warp 0 finds every bin's feature-column span; one chunk when the row fits
        int mode = empty ? 2 : (tables ? 0 : 1), n = 0, maxcol = 1, maxnb = 1;
        if (mode == 0) {
            int glo = INT_MAX, ghi = -1;
            for (int pw = lane; pw < PW; pw += 32) {
                int lo = INT_MAX, hi = -1;
                for (int ix = 0; ix < g.grid_w; ++ix) {
                    const XTap X = xs[pw * g.grid_w + ix];
                    if (X.lo >= 0) { lo = min(lo, X.lo); hi = max(hi, X.hi); }
                }
                blo[pw] = lo; bhi[pw] = hi;
                glo = min(glo, lo); ghi = max(ghi, hi);
            }
            glo = __reduce_min_sync(0xffffffffu, glo);
            ghi = __reduce_max_sync(0xffffffffu, ghi);
            __syncwarp();
            if (ghi < 0) {
                mode = 2;                                   // every x sample lies outside the map
            } else if (PW <= 32 && ghi - glo + 1 <= kSepCells) {
                n = 1; maxcol = ghi - glo + 1; maxnb = PW;
                if (lane == 0) { SepChunk c; c.pa = 0; c.nb = PW; c.ca = glo; c.ncol = maxcol; chunks[0] = c; }
            } else {                                        // greedy chunks: <= kSepCells columns, <= 32 bins each
                int pa = 0;                                 // (computed redundantly by every lane of warp 0)
                while (pa < PW) {
                    int lo = INT_MAX, hi = -1, pb = pa;
                    while (pb < PW && pb - pa < 32) {
                        const int nlo = min(lo, blo[pb]), nhi = max(hi, bhi[pb]);
                        if (nhi >= 0 && nhi - nlo + 1 > kSepCells) break;
                        lo = nlo; hi = nhi; ++pb;
                    }
                    if (pb == pa || n == kSepChunks) { mode = 1; break; }   // a single bin wider than the buffer
                    SepChunk c;
                    c.pa = pa; c.nb = pb - pa; c.ca = hi < 0 ? 0 : lo; c.ncol = hi < 0 ? 0 : hi - lo + 1;
                    if (lane == 0) chunks[n] = c;
                    ++n;
                    maxcol = max(maxcol, c.ncol); maxnb = max(maxnb, c.nb);
                    pa = pb;
                }
            }
        }
        if (lane == 0) {
            s_mode = mode; s_nchunks = n;
            s_rows = max(1, min(min(kSepCells / maxcol, 32 / maxnb), PH));
        }
    }
    __syncthreads();
    const int mode = s_mode;

    if (mode == 2) {   // no sample inside the map: the RoI pools to zeros
        for (int sl = 0; sl < nslab; ++sl) {
            const int c0 = cg0 + sl * CC, cc = min(CC, C - c0);
            OutT* o = oroi + (size_t)c0 * NB;
            for (int e = threadIdx.x; e < cc * NB; e += blockDim.x) o[e] = from_f32<OutT>(0.0f);
        }
        return;
    }
    if (mode == 1) {   // exotic geometry (sampling grids beyond the tables): direct 4-tap evaluation
        for (int sl = 0; sl < nslab; ++sl) {
            const int c0 = cg0 + sl * CC;
            for (int b = warp; b < NB; b += nwarps) {
                const int ph = b / PW, pw = b - ph * PW;
                float acc[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) acc[j] = 0.0f;
                for (int iy = 0; iy < g.grid_h; ++iy) {
                    const Tap Y = make_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H, W * C);
                    if (Y.lo < 0) continue;
                    for (int ix = 0; ix < g.grid_w; ++ix) {
                        const Tap X = make_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W, C);
                        if (X.lo < 0) continue;
                        const float w1 = Y.h * X.h, w2 = Y.h * X.l, w3 = Y.l * X.h, w4 = Y.l * X.l;
#pragma unroll
                        for (int j = 0; j < CPL; ++j) {
                            const int c = c0 + lane + 32 * j;
                            if (c < C) {
                                const float* f = fimg + c;
                                acc[j] += w1 * __ldg(f + Y.lo + X.lo) + w2 * __ldg(f + Y.lo + X.hi) +
                                          w3 * __ldg(f + Y.hi + X.lo) + w4 * __ldg(f + Y.hi + X.hi);
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const int c = c0 + lane + 32 * j;
                    if (c < C) oroi[(size_t)c * NB + b] = from_f32<OutT>(acc[j] * rcount);
                }
            }
        }
        return;
    }

    // ---- separable path ---------------------------------------------------------------------
    const int nchunks = s_nchunks, rows = s_rows;
    const int nrg = (PH + rows - 1) / rows, nunits = nrg * nchunks;
    float* __restrict__ Tw = tsm + (size_t)warp * kSepCells * P;
    const int gh = g.grid_h, gw = g.grid_w;

    for (int item = warp; item < nunits * nslab; item += nwarps) {
        const int sl = item / nunits, unit = item - sl * nunits;
        const int rg = unit / nchunks;
        const SepChunk ch = chunks[unit - rg * nchunks];
        const int ph0 = rg * rows, nr = min(rows, PH - ph0);
        const int c0 = cg0 + sl * CC, cc = min(CC, C - c0);
        const float* __restrict__ fb = fimg + c0 + lane + (size_t)ch.ca * C;

        // phase 1: T[row][col][channel] for the unit (lane = channel)
        for (int r = 0; r < nr; ++r) {
            const Tap* __restrict__ yt = ys + (ph0 + r) * gh;
            float* __restrict__ trow = Tw + r * ch.ncol * P + lane;
            if (cc == CC && ch.ncol >= 8) sep_phase1_row<8, CPL, true>(fb, yt, gh, ch.ncol, C, cc, lane, trow);
            else if (cc == CC && ch.ncol >= 4) sep_phase1_row<4, CPL, true>(fb, yt, gh, ch.ncol, C, cc, lane, trow);
            else sep_phase1_row<4, CPL, false>(fb, yt, gh, ch.ncol, C, cc, lane, trow);
        }
        __syncwarp();

        // phase 2: lane = (channel split, output bin). With <= 16 bins in the unit the two half-warps take
        // the two halves of the channel slab, so wide RoIs (one row per unit) still use 28 lanes.
        const int nact = nr * ch.nb;
        const int nsplit = nact <= 16 ? 2 : 1;
        const int split = nsplit == 2 ? lane >> 4 : 0;
        const int idx = nsplit == 2 ? lane & 15 : lane;
        const bool active = idx < nact;
        const int lr = active ? idx / ch.nb : 0;
        const int pwl = active ? idx - lr * ch.nb : 0;
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, w3 = 0.0f;
        int first = -1, span = 0;
        if (active) {
            const XTap* xt = xs + (ch.pa + pwl) * gw;
            for (int ix = 0; ix < gw; ++ix) {
                const XTap X = xt[ix];
                if (X.lo < 0) continue;
                if (first < 0) first = X.lo;
                const int d = X.lo - first, e = X.hi - first;
                if (d < 0) span = 99;
                span = max(span, e + 1);
                w0 += (d == 0 ? X.h : 0.0f) + (e == 0 ? X.l : 0.0f);
                w1 += (d == 1 ? X.h : 0.0f) + (e == 1 ? X.l : 0.0f);
                w2 += (d == 2 ? X.h : 0.0f) + (e == 2 ? X.l : 0.0f);
                w3 += (d == 3 ? X.h : 0.0f) + (e == 3 ? X.l : 0.0f);
            }
        }
        const int maxspan = __reduce_max_sync(0xffffffffu, span);
        const int cb = first < 0 ? 0 : first - ch.ca;            // first column of the bin inside the chunk
        const int half = (CC / 2) * split;                        // first channel of this lane's share
        const int nch = nsplit == 2 ? max(0, min(cc - half, CC / 2)) : cc;
        const float* __restrict__ tb = Tw + (lr * ch.ncol + cb) * P + half;
        OutT* __restrict__ ob = oroi + (size_t)(c0 + half) * NB + (ph0 + lr) * PW + ch.pa + pwl;
        if (maxspan <= 4) {
            // Tap d of a lane whose bin spans fewer than d+1 columns has weight 0 and re-reads the bin's own
            // last column, so a zero weight never meets a value the reference would not have touched.
            const int last = max(span - 1, 0);
            const int o1 = min(1, last) * P, o2 = min(2, last) * P, o3 = min(3, last) * P;
            w0 *= rcount; w1 *= rcount; w2 *= rcount; w3 *= rcount;
            if (active && first >= 0) {
                if (maxspan <= 2) sep_phase2_loop<OutT, 2, P>(tb, o1, o2, o3, w0, w1, w2, w3, ob, NB, nch);
                else if (maxspan == 3) sep_phase2_loop<OutT, 3, P>(tb, o1, o2, o3, w0, w1, w2, w3, ob, NB, nch);
                else sep_phase2_loop<OutT, 4, P>(tb, o1, o2, o3, w0, w1, w2, w3, ob, NB, nch);
            } else if (active) {   // no valid x sample for this bin
                for (int c = 0; c < nch; ++c) ob[(size_t)c * NB] = from_f32<OutT>(0.0f);
            }
        } else if (maxspan <= 8) {
            // bins spanning 5..8 feature columns (RoIs wider than ~42 cells, e.g. the whole 75-cell map): the
            // same scheme with 8 combined column weights (recomputed here so the common path keeps 4 registers)
            float w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) w[q] = 0.0f;
            if (active) {
                const XTap* xt = xs + (ch.pa + pwl) * gw;
                for (int ix = 0; ix < gw; ++ix) {
                    const XTap X = xt[ix];
                    if (X.lo < 0) continue;
                    const int d = X.lo - first, e = X.hi - first;
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] += (d == q ? X.h : 0.0f) + (e == q ? X.l : 0.0f);
                }
            }
            const int last = max(span - 1, 0);
            int o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { w[q] *= rcount; o[q] = min(q, last) * P; }
            if (active) {
                for (int c = 0; c < nch; ++c) {
                    float acc = first >= 0 ? w[0] * tb[c] : 0.0f;
                    if (first >= 0) {
#pragma unroll
                        for (int q = 1; q < 8; ++q) acc = __fmaf_rn(w[q], tb[o[q] + c], acc);
                    }
                    ob[(size_t)c * NB] = from_f32<OutT>(acc);
                }
            }
        } else if (active) {   // still wider bins (fixed sampling_ratio with huge RoIs): per-sample evaluation
            const XTap* xt = xs + (ch.pa + pwl) * gw;
            const float* __restrict__ trow = Tw + (lr * ch.ncol - ch.ca) * P + half;
            for (int c = 0; c < nch; ++c) {
                float acc = 0.0f;
                for (int ix = 0; ix < gw; ++ix) {
                    const XTap X = xt[ix];
                    if (X.lo >= 0) acc = __fmaf_rn(X.l, trow[X.hi * P + c], __fmaf_rn(X.h, trow[X.lo * P + c], acc));
                }
                ob[(size_t)c * NB] = from_f32<OutT>(acc * rcount);
            }
        }
        __syncwarp();
    }
}

template <typename OutT, int CPL, int CS, int PHT, int PWT>
static int launch_sep(const RoiParams& p, OutT* out, int warps, int slabs, cudaStream_t s) {
    constexpr int CC = 32 * CPL;
    auto kern = roi_align_fwd_sep_kernel<OutT, CPL, CS, PHT, PWT>;
    const size_t smem = (size_t)warps * kSepCells * (CC + 1) * sizeof(float);
    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int cgroups = (int)ceil_div(p.C, CC * slabs);
    kern<<<(unsigned)(p.K * cgroups), 32 * warps, smem, s>>>(p, out, cgroups, slabs);
    return check_launch("roi_align_fwd_sep_kernel");
}

// ------------------------------------------------------------------------------------------------
// backward: the adjoint of the separable forward
//   grad_in[y][x][c] += 1/count * sum_{ph,pw} Wy[ph][y] * Wx[pw][x] * g[c][ph][pw]
// One warp owns a unit = R (<= 2) output rows x all PW bins x a 32*CPL-channel slab:
//   stage   (lane = bin): the unit's grad_out values, coalesced from [K,C,PH,PW], into a per-warp shared
//           tile G[bin][channel] with an odd pitch (conflict-free both ways);
//   walk    (lane = channel): one pass along x for both rows at once (they share the x taps) with a
//           two-column register window accumulating sum_pw Wx * g;
//   flush   when a column leaves the window: for every feature row y touched by the unit (the y taps of
//           its rows merged into a table once per CTA), ONE fp32 RED of Wy-weighted window values -
//           rows of the unit that share a feature row share the atomic.
// ------------------------------------------------------------------------------------------------
constexpr int kBwdYEnt = 16;    // merged y-table entries per unit (7x7 outputs of large RoIs reach grid_h = 5)
constexpr int kBwdUnits = 8;    // units per RoI handled by the tables (PH <= 16 with two rows per unit)
constexpr int kBwdCols = 128;   // feature columns of one RoI handled by the column program
constexpr int kBwdEnt = 192;    // (bin, weight) entries of the column program

struct YEnt { int off; float w0, w1; int pad; };   // feature-row offset (y*W*C), weights of unit rows 0 and 1
struct CEnt { int bin; float w; };                 // one bin's combined x weight on a feature column

// walk + flush for one unit (lane = channel). FULL: the whole 32*CPL-channel slab is inside C.
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// walk + flush for one unit. FULL: the whole channel slab is inside C.
// VEC = false: lane owns channels lane + 32*j of the slab (tile pitch 32*CPL + 1, scalar fp32 REDs).
// VEC = true (CPL == 2, even C): lane owns the ADJACENT channels 2*lane, 2*lane + 1 (tile pitch 66): the tile is
// read with 64-bit LDS and every flush is ONE vector RED (red.global.add.v2.f32) per lane instead of two - the
// flush, not the atomic unit, is what the backward spends most of its time issuing.
template <int CPL, bool FULL, bool VEC>
__device__ __forceinline__ void bwd_walk(const float* __restrict__ Gw, float* __restrict__ gb, const int C, const int PW,
                                         const int nr, const int cc, const int lane, const int cmin, const int ncols,
                                         const int* __restrict__ colstart, const CEnt* __restrict__ cent,
                                         const YEnt* __restrict__ yt, const int ne) {
    constexpr int P = VEC ? 32 * CPL + 2 : 32 * CPL + 1;
    if (VEC) {
        const float* __restrict__ G0 = Gw + 2 * lane;
        const float* __restrict__ G1 = Gw + (nr == 2 ? PW * P : 0) + 2 * lane;
        const float s1 = nr == 2 ? 1.0f : 0.0f;
        const bool ok0 = FULL || 2 * lane < cc, ok1 = FULL || 2 * lane + 1 < cc;
        float* __restrict__ gl = gb + 2 * lane;
        for (int ci = 0; ci < ncols; ++ci) {
            const int e0 = colstart[ci], e1 = colstart[ci + 1];
            if (e0 == e1) continue;
            float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
            for (int e = e0; e < e1; ++e) {
                const CEnt E = cent[e];
                const float2 g0 = *reinterpret_cast<const float2*>(G0 + E.bin * P);
                const float2 g1 = *reinterpret_cast<const float2*>(G1 + E.bin * P);
                a0.x = __fmaf_rn(E.w, g0.x, a0.x); a0.y = __fmaf_rn(E.w, g0.y, a0.y);
                a1.x = __fmaf_rn(E.w, g1.x, a1.x); a1.y = __fmaf_rn(E.w, g1.y, a1.y);
            }
            float* __restrict__ gc = gl + (size_t)(cmin + ci) * C;
            for (int e = 0; e < ne; ++e) {
                const YEnt Y = yt[e];
                const float w1 = Y.w1 * s1;
                const float vx = __fmaf_rn(w1, a1.x, Y.w0 * a0.x), vy = __fmaf_rn(w1, a1.y, Y.w0 * a0.y);
                if (ok1) red_add_v2(gc + Y.off, vx, vy);
                else if (ok0) atomicAdd(gc + Y.off, vx);
            }
        }
        return;
    }
    const float* __restrict__ G0 = Gw + lane;
    const float* __restrict__ G1 = Gw + (nr == 2 ? PW * P : 0) + lane;
    const float s1 = nr == 2 ? 1.0f : 0.0f;      // a single-row unit has no second row
    float* __restrict__ gl = gb + lane;
    for (int ci = 0; ci < ncols; ++ci) {
        const int e0 = colstart[ci], e1 = colstart[ci + 1];
        if (e0 == e1) continue;
        float a0[CPL], a1[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) { a0[j] = 0.0f; a1[j] = 0.0f; }
        for (int e = e0; e < e1; ++e) {
            const CEnt E = cent[e];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const bool ok = FULL || lane + 32 * j < cc;
                const float g0 = ok ? G0[E.bin * P + 32 * j] : 0.0f;
                const float g1 = ok ? G1[E.bin * P + 32 * j] : 0.0f;
                a0[j] = __fmaf_rn(E.w, g0, a0[j]);
                a1[j] = __fmaf_rn(E.w, g1, a1[j]);
            }
        }
        float* __restrict__ gc = gl + (size_t)(cmin + ci) * C;
        for (int e = 0; e < ne; ++e) {
            const YEnt Y = yt[e];
            const float w1 = Y.w1 * s1;
#pragma unroll
            for (int j = 0; j < CPL; ++j)
                if (FULL || lane + 32 * j < cc) atomicAdd(gc + Y.off + 32 * j, __fmaf_rn(w1, a1[j], Y.w0 * a0[j]));
        }
    }
}

template <typename GT, int CPL, int CS, int PHT, int PWT, int SB, bool VEC>
__global__ void __launch_bounds__(224, 4)
roi_align_bwd_sep_kernel(const RoiParams p, const GT* __restrict__ go, const int cgroups, const int slabs,
                         const int trows) {
    constexpr int CC = 32 * CPL, P = VEC ? CC + 2 : CC + 1;
    extern __shared__ __align__(16) float gsm[];   // [nwarps][32][P]; its head doubles as the tap tables below
    // the x / y tap tables are only needed while the column program and the y tables are built: they live in
    // the (not yet used) per-warp tile region, which keeps the static footprint at ~4 KB -> 4 CTAs per SM
    XTap* xs = reinterpret_cast<XTap*>(gsm);
    Tap* ys = reinterpret_cast<Tap*>(gsm) + kSepTap;
    __shared__ YEnt ytab[kBwdUnits][kBwdYEnt];
    __shared__ int ycnt[kBwdUnits];
    __shared__ int colstart[kBwdCols + 1];
    __shared__ CEnt cent[kBwdEnt];
    __shared__ int s_mode, s_cmin, s_cmax;         // mode 0: separable, 1: direct

    const int kk = blockIdx.x / cgroups;
    if (p.k_dev && kk >= __ldg(p.k_dev)) return;   // capacity launch: RoI beyond the live count
    const int k = p.perm ? __ldg(p.perm + kk) : kk;   // launch order / RoI subset (coin_roi_launch_order, coin_roi_split_by_area)
    const int cg0 = (blockIdx.x - kk * cgroups) * (CC * slabs);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
    const coin_level_t L = p.lv[lvl];
    const int H = L.H, W = L.W;
    const int C = CS ? CS : p.C, PH = PHT ? PHT : p.PH, PW = PWT ? PWT : p.PW, NB = PH * PW;
    const RoiGeom g = roi_geometry(p.rois + 5 * (size_t)k, L.spatial_scale, PH, PW, p.sampling_ratio, p.aligned);
    float* __restrict__ gimg = const_cast<float*>(L.feat_nhwc) + (size_t)g.batch * H * W * C;
    const float rcount = 1.0f / g.count;
    const int nslab = min(slabs, (int)((C - cg0 + CC - 1) / CC));
    const GT* __restrict__ groi = go + (size_t)k * C * NB;
    if (g.grid_h <= 0 || g.grid_w <= 0) return;    // no samples: no gradient
    const int gh = g.grid_h, gw = g.grid_w;
    const int R = PW <= 16 && PH >= 2 ? 2 : 1;     // rows per unit (R*PW bins <= 32 lanes)
    const int nunits = (PH + R - 1) / R;
    const bool tables = PW <= 32 && (long long)PW * gw <= kSepTap && (long long)PH * gh <= kSepTap && nunits <= kBwdUnits;
    if (threadIdx.x == 0) { s_mode = tables ? 0 : 1; s_cmin = INT_MAX; s_cmax = -1; }
    if (tables) {
        const int nx = PW * gw, ny = PH * gh;
        for (int s = threadIdx.x; s < nx; s += blockDim.x) {
            const int pw = s / gw;
            xs[s] = make_xtap(g.start_w, g.bin_w, pw, s - pw * gw, gw, W);
        }
        for (int s = blockDim.x - 1 - threadIdx.x; s < ny; s += blockDim.x) {
            const int ph = s / gh;
            ys[s] = make_tap(g.start_h, g.bin_h, ph, s - ph * gh, gh, H, W * C);
        }
    }
    __syncthreads();
    if (tables) {
        // merged y table of every unit (one thread per unit, from the high thread ids)
        for (int u = blockDim.x - 1 - threadIdx.x; u < nunits; u += blockDim.x) {
            int n = 0;
            bool overflow = false;
            for (int r = 0; r < R && u * R + r < PH; ++r)
                for (int iy = 0; iy < gh; ++iy) {
                    const Tap Y = ys[(u * R + r) * gh + iy];
                    if (Y.lo < 0) continue;
                    for (int t = 0; t < 2; ++t) {
                        const int off = t ? Y.hi : Y.lo;
                        const float w = (t ? Y.l : Y.h) * rcount;
                        int e = 0;
                        while (e < n && ytab[u][e].off != off) ++e;
                        if (e == n) {
                            if (n == kBwdYEnt) { overflow = true; break; }
                            ytab[u][e].off = off; ytab[u][e].w0 = 0.0f; ytab[u][e].w1 = 0.0f; ytab[u][e].pad = 0;
                            ++n;
                        }
                        if (r == 0) ytab[u][e].w0 += w; else ytab[u][e].w1 += w;
                    }
                }
            ycnt[u] = n;
            if (overflow) s_mode = 1;
        }
        // feature-column range of the RoI (warp 0)
        if (warp == 0) {
            int lo = INT_MAX, hi = -1;
            for (int s = lane; s < PW * gw; s += 32) {
                const XTap X = xs[s];
                if (X.lo >= 0) { lo = min(lo, X.lo); hi = max(hi, X.hi); }
            }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0) { s_cmin = lo; s_cmax = hi; }
        }
    }
    __syncthreads();
    const int cmin = s_cmin, ncols = s_cmax < 0 ? 0 : s_cmax - s_cmin + 1;
    if (tables && ncols == 0) return;              // every x sample lies outside the map
    if (tables && ncols > kBwdCols) { if (threadIdx.x == 0) s_mode = 1; }
    else if (tables) {
        // column program, pass 1: entries per column = bins with a non-empty weight on it
        for (int ci = threadIdx.x; ci < ncols; ci += blockDim.x) {
            const int col = cmin + ci;
            int n = 0;
            for (int pw = 0; pw < PW; ++pw) {
                bool hit = false;
                for (int ix = 0; ix < gw; ++ix) {
                    const XTap X = xs[pw * gw + ix];
                    hit |= X.lo >= 0 && (X.lo == col || X.hi == col);
                }
                n += hit;
            }
            colstart[ci + 1] = n;
        }
    }
    __syncthreads();
    if (s_mode == 0) {
        if (threadIdx.x == 0) {   // exclusive prefix over <= kBwdCols columns
            int acc = 0;
            colstart[0] = 0;
            for (int ci = 0; ci < ncols; ++ci) { acc += colstart[ci + 1]; colstart[ci + 1] = acc; }
            if (acc > kBwdEnt) s_mode = 1;
        }
    }
    __syncthreads();
    if (s_mode == 0) {
        for (int ci = threadIdx.x; ci < ncols; ci += blockDim.x) {   // pass 2: fill (bin, combined weight)
            const int col = cmin + ci;
            int at = colstart[ci];
            for (int pw = 0; pw < PW; ++pw) {
                float w = 0.0f;
                bool hit = false;
                for (int ix = 0; ix < gw; ++ix) {
                    const XTap X = xs[pw * gw + ix];
                    if (X.lo < 0) continue;
                    if (X.lo == col) { w += X.h; hit = true; }
                    if (X.hi == col) { w += X.l; hit = true; }
                }
                if (hit) { cent[at].bin = pw; cent[at].w = w; ++at; }
            }
        }
    }
    __syncthreads();

    if (s_mode == 1) {   // exotic geometry: direct 4-tap scatter
        for (int sl = 0; sl < nslab; ++sl) {
            const int c0 = cg0 + sl * CC;
            for (int b = warp; b < NB; b += nwarps) {
                const int ph = b / PW, pw = b - ph * PW;
                float gv[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const int c = c0 + lane + 32 * j;
                    gv[j] = c < C ? to_f32(groi[(size_t)c * NB + b]) * rcount : 0.0f;
                }
                for (int iy = 0; iy < gh; ++iy) {
                    const Tap Y = make_tap(g.start_h, g.bin_h, ph, iy, gh, H, W * C);
                    if (Y.lo < 0) continue;
                    for (int ix = 0; ix < gw; ++ix) {
                        const Tap X = make_tap(g.start_w, g.bin_w, pw, ix, gw, W, C);
                        if (X.lo < 0) continue;
                        const float w1 = Y.h * X.h, w2 = Y.h * X.l, w3 = Y.l * X.h, w4 = Y.l * X.l;
#pragma unroll
                        for (int j = 0; j < CPL; ++j) {
                            const int c = c0 + lane + 32 * j;
                            if (c < C) {
                                float* f = gimg + c;
                                atomicAdd(f + Y.lo + X.lo, w1 * gv[j]);
                                atomicAdd(f + Y.lo + X.hi, w2 * gv[j]);
                                atomicAdd(f + Y.hi + X.lo, w3 * gv[j]);
                                atomicAdd(f + Y.hi + X.hi, w4 * gv[j]);
                            }
                        }
                    }
                }
            }
        }
        return;
    }

    float* __restrict__ Gw = gsm + (size_t)warp * trows * P;   // trows = bins per unit (R * PW <= 32)
    const bool prefetch = p.flags & 1;
    for (int item = warp; item < nunits * nslab; item += nwarps) {
        const int sl = item / nunits, unit = item - sl * nunits;
        const int ph0 = unit * R, nr = min(R, PH - ph0);
        const int c0 = cg0 + sl * CC, cc = min(CC, C - c0);
        const int nbu = nr * PW;
        const int ne = ycnt[unit];
        if (ne == 0) continue;                     // every y sample of the unit lies outside the map
        // ---- stage: lane = bin, coalesced rows of grad_out -> G[bin][channel]
        if (lane < nbu) {
            const GT* __restrict__ gp = groi + (size_t)c0 * NB + ph0 * PW + lane;
            float* __restrict__ gw_ = Gw + lane * P;
            int c = 0;
            for (; c + SB <= cc; c += SB) {   // SB loads in flight per lane
                float v[SB];
#pragma unroll
                for (int q = 0; q < SB; ++q) v[q] = to_f32(__ldg(gp + (size_t)(c + q) * NB));
#pragma unroll
                for (int q = 0; q < SB; ++q) gw_[c + q] = v[q];
            }
            for (; c < cc; ++c) gw_[c] = to_f32(__ldg(gp + (size_t)c * NB));
        }
        __syncwarp();
        // L2 prefetch of the NEXT item's grad_out rows (this warp's next unit/slab): the rows stream from HBM
        // exactly once, so without it every staging load pays the full DRAM latency
        {
            const int nitem = item + nwarps;
            if (prefetch && nitem < nunits * nslab) {
                const int nsl = nitem / nunits, nun = nitem - nsl * nunits;
                const int nph0 = nun * R, nnr = min(R, PH - nph0);
                const int nc0 = cg0 + nsl * CC, ncc = min(CC, C - nc0);
                const int bytes = nnr * PW * (int)sizeof(GT);          // contiguous bytes per channel row
                const int segs = (bytes + 31) / 32 + 1;                 // 32-byte sectors (+1 for misalignment)
                const char* base = reinterpret_cast<const char*>(groi + (size_t)nc0 * NB + nph0 * PW);
                for (int e = lane; e < ncc * segs; e += 32) {
                    const int ch = e / segs, sg = e - ch * segs;
                    const int off = min(sg * 32, bytes - 1);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)ch * NB * sizeof(GT) + off));
                }
            }
        }
        float* __restrict__ gb = gimg + c0;
        if (cc == CC) bwd_walk<CPL, true, VEC>(Gw, gb, C, PW, nr, cc, lane, cmin, ncols, colstart, cent, ytab[unit], ne);
        else bwd_walk<CPL, false, VEC>(Gw, gb, C, PW, nr, cc, lane, cmin, ncols, colstart, cent, ytab[unit], ne);
        __syncwarp();
    }
}

template <typename GT, int CPL, int CS, int PHT, int PWT, int SB, bool VEC = false>
static int launch_bwd_sep(const RoiParams& p, const GT* go, int warps, int slabs, cudaStream_t s) {
    constexpr int CC = 32 * CPL;
    auto kern = roi_align_bwd_sep_kernel<GT, CPL, CS, PHT, PWT, SB, VEC>;
    const int R = p.PW <= 16 && p.PH >= 2 ? 2 : 1;
    const int trows = std::min(32, R * p.PW);
    const size_t smem = std::max<size_t>((size_t)warps * trows * (CC + (VEC ? 2 : 1)) * sizeof(float), 2 * kSepTap * 16);
    if (smem > 24 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int cgroups = (int)ceil_div(p.C, CC * slabs);
    kern<<<(unsigned)(p.K * cgroups), 32 * warps, smem, s>>>(p, go, cgroups, slabs, trows);
    return check_launch("roi_align_bwd_sep_kernel");
}

template <typename GT>
static int dispatch_bwd_sep(const RoiParams& p, const GT* g, int cpl, int warps, int slabs, cudaStream_t s) {
    const int sb = sep_env("COIN_ROI_BWD_SB", 16);
    // vector REDs need 8-byte aligned channel pairs: even C (slab starts are multiples of 64)
    const bool vec = cpl == 2 && p.C % 2 == 0 && sep_env("COIN_ROI_BWD_VEC", 1) != 0;
    if (p.C == 1024 && p.PH == 14 && p.PW == 14) {
        if (cpl == 1) return launch_bwd_sep<GT, 1, 1024, 14, 14, 16>(p, g, warps, slabs, s);
        if (sb == 32) return launch_bwd_sep<GT, 2, 1024, 14, 14, 32>(p, g, warps, slabs, s);
        if (vec) return launch_bwd_sep<GT, 2, 1024, 14, 14, 16, true>(p, g, warps, slabs, s);
        return launch_bwd_sep<GT, 2, 1024, 14, 14, 16>(p, g, warps, slabs, s);
    }
    if (cpl == 1) return launch_bwd_sep<GT, 1, 0, 0, 0, 16>(p, g, warps, slabs, s);
    if (p.C == 1024 && p.PH == 7 && p.PW == 7) {
        if (vec) return launch_bwd_sep<GT, 2, 1024, 7, 7, 16, true>(p, g, warps, slabs, s);
        return launch_bwd_sep<GT, 2, 1024, 7, 7, 16>(p, g, warps, slabs, s);
    }
    if (vec) return launch_bwd_sep<GT, 2, 0, 0, 0, 16, true>(p, g, warps, slabs, s);
    return launch_bwd_sep<GT, 2, 0, 0, 0, 16>(p, g, warps, slabs, s);
}

template <typename OutT>
static int dispatch_sep(const RoiParams& p, OutT* out, int cpl, int warps, int slabs, cudaStream_t s) {
    if (cpl == 1) return launch_sep<OutT, 1, 0, 0, 0>(p, out, warps, slabs, s);
    const bool spec = sep_env("COIN_ROI_SEP_GENERIC", 0) == 0;
    if (spec && p.C == 1024 && p.PH == 14 && p.PW == 14) return launch_sep<OutT, 2, 1024, 14, 14>(p, out, warps, slabs, s);
    if (spec && p.C == 1024 && p.PH == 7 && p.PW == 7) return launch_sep<OutT, 2, 1024, 7, 7>(p, out, warps, slabs, s);
    if (spec && p.C == 256 && p.PH == 7 && p.PW == 7) return launch_sep<OutT, 2, 256, 7, 7>(p, out, warps, slabs, s);
    return launch_sep<OutT, 2, 0, 0, 0>(p, out, warps, slabs, s);
}

int launch_roi_align_fwd_sep(const RoiParams& p, void* out, int out_dtype, cudaStream_t s) {
    int cpl = sep_env("COIN_ROI_SEP_CPL", 2);
    if (cpl != 1 && cpl != 2) cpl = 2;
    if (p.C <= 32) cpl = 1;
    const int cc = 32 * cpl;
    // one warp per unit of the common case (two 14-bin rows per unit, or four 7-bin rows)
    const int rows = std::max(1, std::min(32 / std::min(p.PW, 32), p.PH));
    // small outputs (7x7): little work per unit, so more warps and fewer slabs per CTA (measured)
    const bool small = p.PH * p.PW <= 64;
    int warps = sep_env("COIN_ROI_SEP_WARPS", small ? 8 : (int)std::min<int64_t>(8, std::max<int64_t>(ceil_div(p.PH, rows), 4)));
    warps = std::max(1, std::min(warps, 8));
    // few RoIs (e.g. the private-box call of the step): one slab per CTA, so a single very large RoI is spread
    // over C/64 CTAs instead of leaving a long tail on C/256 of them
    int slabs = sep_env("COIN_ROI_SEP_SLABS", small ? 2 : (p.K < 1024 ? 1 : 4));
    slabs = std::max(1, std::min(slabs, (int)ceil_div(p.C, cc)));
    if (out_dtype == COIN_F32) return dispatch_sep<float>(p, static_cast<float*>(out), cpl, warps, slabs, s);
    return dispatch_sep<__half>(p, static_cast<__half*>(out), cpl, warps, slabs, s);
}


int launch_roi_align_bwd_sep(const RoiParams& p_in, const void* grad_out, int grad_dtype, cudaStream_t s) {
    RoiParams p = p_in;
    p.flags = sep_env("COIN_ROI_BWD_PREFETCH", 1) ? 1 : 0;
    const int cpl = p.C <= 32 ? 1 : (sep_env("COIN_ROI_BWD_CPL", 2) == 1 ? 1 : 2);
    const int R = p.PW <= 16 && p.PH >= 2 ? 2 : 1;
    int warps = sep_env("COIN_ROI_BWD_WARPS", (int)std::min<int64_t>(7, std::max<int64_t>(ceil_div(p.PH, R), 4)));
    warps = std::max(1, std::min(warps, 7));   // the kernel is compiled for <= 224 threads
    int slabs = sep_env("COIN_ROI_BWD_SLABS", p.PH * p.PW <= 64 ? 4 : 8);
    slabs = std::max(1, std::min(slabs, (int)ceil_div(p.C, 32 * cpl)));
    if (grad_dtype == COIN_F32) return dispatch_bwd_sep<float>(p, static_cast<const float*>(grad_out), cpl, warps, slabs, s);
    return dispatch_bwd_sep<__half>(p, static_cast<const __half*>(grad_out), cpl, warps, slabs, s);
}

}  // namespace coin
