/*
 * oracle/scalar_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, plain C).
 *
 * Scalar restatement of the third-party arithmetic the COIN RoI path reaches
 * (the reference itself ships no native code):
 *
 *   roi_align fwd/bwd : torchvision csrc/ops/cpu/roi_align_kernel.cpp  [tv, restated from the
 *                       published algorithm; formula also visible in the installed
 *                       torchvision/ops/roi_align.py::_roi_align]. Reached from the reference at
 *                       coin/modeling/roi_heads/clip_roi_heads.py:62,173 (ROIPooler -> ROIAlign).
 *   nms               : torchvision csrc/ops/cpu/nms_kernel.cpp [tv, restated]. Reached from
 *                       coin/modeling/roi_heads/fast_rcnn.py:164, coin/layers/nms.py:207.
 *   pairwise_iou      : detectron2 0.5 structures/boxes.py [d2, restated]. Reached from
 *                       coin/engine/trainer.py:364,373, coin/utils/util.py:468,
 *                       coin/modeling/roi_heads/clip_roi_heads.py:301,311,353,
 *                       coin/modeling/proposal_generator/rpn.py:159,169,212.
 *
 * Pinned (tests/test_oracle_cpu.py) against torchvision 0.26 CPU ops run in this image and the
 * committed fixtures in tests/golden/. Nothing under coin_b200/ may link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it, as the checker.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp). -ffp-contract=off keeps the
 * multiply/add sequence un-fused, which is what the x86-64 torchvision wheel executes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* ROIAlign                                                                                    */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    float start_w, start_h, bin_w, bin_h;
    int grid_h, grid_w;
    float count;
    int batch;
} roi_geom_t;

static void roi_geometry(const float* roi, float spatial_scale, int PH, int PW, int sampling_ratio,
                         int aligned, roi_geom_t* g) {
    g->batch = (int)roi[0];
    const float offset = aligned ? 0.5f : 0.0f;
    g->start_w = roi[1] * spatial_scale - offset;
    g->start_h = roi[2] * spatial_scale - offset;
    const float end_w = roi[3] * spatial_scale - offset;
    const float end_h = roi[4] * spatial_scale - offset;
    float roi_w = end_w - g->start_w;
    float roi_h = end_h - g->start_h;
    if (!aligned) {
        roi_w = fmaxf(roi_w, 1.0f);
        roi_h = fmaxf(roi_h, 1.0f);
    }
    g->bin_h = roi_h / (float)PH;
    g->bin_w = roi_w / (float)PW;
    g->grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / (float)PH);
    g->grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / (float)PW);
    const int cnt = g->grid_h * g->grid_w;
    g->count = (float)(cnt > 1 ? cnt : 1);
}

typedef struct {
    int p1, p2, p3, p4; /* flat offsets inside one H*W plane; -1 = sample outside the map */
    float w1, w2, w3, w4;
} tap_t;

static void bilinear_taps(int H, int W, float y, float x, tap_t* t) {
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
        t->p1 = t->p2 = t->p3 = t->p4 = -1;
        t->w1 = t->w2 = t->w3 = t->w4 = 0.0f;
        return;
    }
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int y_low = (int)y, x_low = (int)x, y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else { y_high = y_low + 1; }
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else { x_high = x_low + 1; }
    const float ly = y - (float)y_low, lx = x - (float)x_low;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    t->w1 = hy * hx; t->w2 = hy * lx; t->w3 = ly * hx; t->w4 = ly * lx;
    t->p1 = y_low * W + x_low;  t->p2 = y_low * W + x_high;
    t->p3 = y_high * W + x_low; t->p4 = y_high * W + x_high;
}

static inline float sample_y(const roi_geom_t* g, int ph, int iy) {
    return g->start_h + (float)ph * g->bin_h + ((float)iy + 0.5f) * g->bin_h / (float)g->grid_h;
}
static inline float sample_x(const roi_geom_t* g, int pw, int ix) {
    return g->start_w + (float)pw * g->bin_w + ((float)ix + 0.5f) * g->bin_w / (float)g->grid_w;
}

/* in: [N,C,H,W] fp32 contiguous; rois: [K,5] (batch,x1,y1,x2,y2); out: [K,C,PH,PW]. */
int orc_roi_align_fwd(const float* in, const float* rois, float* out, int N, int C, int H, int W,
                      int K, int PH, int PW, float spatial_scale, int sampling_ratio, int aligned) {
    (void)N;
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < K; ++k) {
        roi_geom_t g;
        roi_geometry(rois + 5 * k, spatial_scale, PH, PW, sampling_ratio, aligned, &g);
        const size_t ntap = (size_t)PH * PW * (size_t)(g.grid_h > 0 ? g.grid_h : 0) *
                            (size_t)(g.grid_w > 0 ? g.grid_w : 0);
        tap_t* taps = (tap_t*)malloc((ntap ? ntap : 1) * sizeof(tap_t));
        size_t idx = 0;
        for (int ph = 0; ph < PH; ++ph)
            for (int pw = 0; pw < PW; ++pw)
                for (int iy = 0; iy < g.grid_h; ++iy)
                    for (int ix = 0; ix < g.grid_w; ++ix)
                        bilinear_taps(H, W, sample_y(&g, ph, iy), sample_x(&g, pw, ix), &taps[idx++]);
        for (int c = 0; c < C; ++c) {
            const float* plane = in + ((size_t)g.batch * C + c) * (size_t)H * W;
            float* o = out + ((size_t)k * C + c) * (size_t)PH * PW;
            idx = 0;
            for (int b = 0; b < PH * PW; ++b) {
                float acc = 0.0f;
                for (int s = 0; s < g.grid_h * g.grid_w; ++s) {
                    const tap_t* t = &taps[idx++];
                    if (t->p1 < 0) continue; /* adds exactly +0 in the original; skipping is identical */
                    acc += t->w1 * plane[t->p1] + t->w2 * plane[t->p2] + t->w3 * plane[t->p3] +
                           t->w4 * plane[t->p4];
                }
                o[b] = acc / g.count;
            }
        }
        free(taps);
    }
    return 0;
}

/* grad_out: [K,C,PH,PW]; grad_in: [N,C,H,W] (zeroed here). Sequential accumulation order
 * (k, c, ph, pw, iy, ix), i.e. the CPU kernel's; the CUDA kernels use atomics, so parity on
 * this op is a tolerance, not bit equality. Not parallelised over k for that reason; the
 * channel loop is (channels never alias). */
int orc_roi_align_bwd(const float* grad_out, const float* rois, float* grad_in, int N, int C, int H,
                      int W, int K, int PH, int PW, float spatial_scale, int sampling_ratio,
                      int aligned) {
    memset(grad_in, 0, (size_t)N * C * H * W * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; ++c) {
        for (int k = 0; k < K; ++k) {
            roi_geom_t g;
            roi_geometry(rois + 5 * k, spatial_scale, PH, PW, sampling_ratio, aligned, &g);
            float* plane = grad_in + ((size_t)g.batch * C + c) * (size_t)H * W;
            const float* go = grad_out + ((size_t)k * C + c) * (size_t)PH * PW;
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    const float gbin = go[ph * PW + pw];
                    for (int iy = 0; iy < g.grid_h; ++iy)
                        for (int ix = 0; ix < g.grid_w; ++ix) {
                            tap_t t;
                            bilinear_taps(H, W, sample_y(&g, ph, iy), sample_x(&g, pw, ix), &t);
                            if (t.p1 < 0) continue;
                            plane[t.p1] += gbin * t.w1 / g.count;
                            plane[t.p2] += gbin * t.w2 / g.count;
                            plane[t.p3] += gbin * t.w3 / g.count;
                            plane[t.p4] += gbin * t.w4 / g.count;
                        }
                }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* pairwise IoU (detectron2 convention: 0 where the intersection is empty)                     */
/* ------------------------------------------------------------------------------------------ */
int orc_pairwise_iou(const float* b1, int n, const float* b2, int m, float* out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        const float* a = b1 + 4 * i;
        const float area_a = (a[2] - a[0]) * (a[3] - a[1]);
        for (int j = 0; j < m; ++j) {
            const float* b = b2 + 4 * j;
            const float area_b = (b[2] - b[0]) * (b[3] - b[1]);
            float w = fminf(a[2], b[2]) - fmaxf(a[0], b[0]);
            float h = fminf(a[3], b[3]) - fmaxf(a[1], b[1]);
            w = w > 0.0f ? w : 0.0f;
            h = h > 0.0f ? h : 0.0f;
            const float inter = w * h;
            out[(size_t)i * m + j] = inter > 0.0f ? inter / (area_a + area_b - inter) : 0.0f;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* greedy NMS: stable descending sort, suppress when IoU > thr (compared in double, as the     */
/* torchvision CPU kernel compares a float IoU with its double threshold argument).            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float s; int64_t i; } sc_t;

static void merge_sort_desc(sc_t* a, sc_t* tmp, int64_t n) {
    if (n < 2) return;
    const int64_t h = n / 2;
    merge_sort_desc(a, tmp, h);
    merge_sort_desc(a + h, tmp, n - h);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].s > a[i].s) ? a[j++] : a[i++]; /* ties: left first */
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, (size_t)n * sizeof(sc_t));
}

int orc_nms(const float* boxes, const float* scores, int64_t n, double thr, int64_t* keep,
            int64_t* nkeep) {
    sc_t* ord = (sc_t*)malloc((size_t)(n ? n : 1) * sizeof(sc_t));
    sc_t* tmp = (sc_t*)malloc((size_t)(n ? n : 1) * sizeof(sc_t));
    float* area = (float*)malloc((size_t)(n ? n : 1) * sizeof(float));
    unsigned char* dead = (unsigned char*)calloc((size_t)(n ? n : 1), 1);
    for (int64_t i = 0; i < n; ++i) {
        ord[i].s = scores[i]; ord[i].i = i;
        area[i] = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    merge_sort_desc(ord, tmp, n);
    int64_t nk = 0;
    for (int64_t a = 0; a < n; ++a) {
        const int64_t i = ord[a].i;
        if (dead[i]) continue;
        keep[nk++] = i;
        const float* bi = boxes + 4 * i;
        for (int64_t b = a + 1; b < n; ++b) {
            const int64_t j = ord[b].i;
            if (dead[j]) continue;
            const float* bj = boxes + 4 * j;
            const float xx1 = fmaxf(bi[0], bj[0]), yy1 = fmaxf(bi[1], bj[1]);
            const float xx2 = fminf(bi[2], bj[2]), yy2 = fminf(bi[3], bj[3]);
            const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
            const float inter = w * h;
            const float ovr = inter / (area[i] + area[j] - inter);
            if ((double)ovr > thr) dead[j] = 1;
        }
    }
    *nkeep = nk;
    free(ord); free(tmp); free(area); free(dead);
    return 0;
}
