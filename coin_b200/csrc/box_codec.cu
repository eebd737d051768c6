// box_codec.cu -- Box2BoxTransform decode / encode, Boxes.clip, Boxes.scale + flip.
//
// Replaces ~20 ATen elementwise launches per apply_deltas call (detectron2 box_regression.py) as
// reached from coin/modeling/roi_heads/fast_rcnn.py:297,619-622,691,729 and the d2 RPN decode
// (<- coin/modeling/proposal_generator/rpn.py:113); Boxes.clip at fast_rcnn.py:145-147; the
// scale + flip of coin/engine/base.py:80-126. Pure streaming kernels: one float4 box per thread,
// 48 B of traffic per class-agnostic RoI. Multiply/add stay un-fused (-fmad=false) so that the
// decoded coordinates follow the oracle's operation order.
#include "common.cuh"

namespace coin {

__global__ void apply_deltas_kernel(const float4* __restrict__ deltas, const float4* __restrict__ boxes,
                                    float4* __restrict__ out, int64_t total, int kreg, float wx, float wy,
                                    float ww, float wh, float scale_clamp, int clip, float clip_h, float clip_w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float4 b = __ldg(boxes + i / kreg);
    const float4 d = __ldg(deltas + i);
    const float w = b.z - b.x, h = b.w - b.y;
    const float cx = b.x + 0.5f * w, cy = b.y + 0.5f * h;
    const float dx = d.x / wx, dy = d.y / wy;
    const float dw = fminf(d.z / ww, scale_clamp), dh = fminf(d.w / wh, scale_clamp);
    const float pcx = dx * w + cx, pcy = dy * h + cy;
    const float pw = expf(dw) * w, ph = expf(dh) * h;
    float4 o = make_float4(pcx - 0.5f * pw, pcy - 0.5f * ph, pcx + 0.5f * pw, pcy + 0.5f * ph);
    if (clip) {
        o.x = fminf(fmaxf(o.x, 0.0f), clip_w);
        o.y = fminf(fmaxf(o.y, 0.0f), clip_h);
        o.z = fminf(fmaxf(o.z, 0.0f), clip_w);
        o.w = fminf(fmaxf(o.w, 0.0f), clip_h);
    }
    out[i] = o;
}

__global__ void get_deltas_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt,
                                  float4* __restrict__ out, int64_t n, float wx, float wy, float ww, float wh,
                                  int32_t* __restrict__ invalid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = __ldg(src + i), t = __ldg(tgt + i);
    const float sw = s.z - s.x, sh = s.w - s.y;
    const float scx = s.x + 0.5f * sw, scy = s.y + 0.5f * sh;
    const float tw = t.z - t.x, th = t.w - t.y;
    const float tcx = t.x + 0.5f * tw, tcy = t.y + 0.5f * th;
    out[i] = make_float4(wx * (tcx - scx) / sw, wy * (tcy - scy) / sh, ww * logf(tw / sw), wh * logf(th / sh));
    if (invalid && !(sw > 0.0f)) *invalid = 1;
}

__global__ void clip_kernel(float4* __restrict__ boxes, int64_t n, float h, float w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = boxes[i];
    b.x = fminf(fmaxf(b.x, 0.0f), w);
    b.y = fminf(fmaxf(b.y, 0.0f), h);
    b.z = fminf(fmaxf(b.z, 0.0f), w);
    b.w = fminf(fmaxf(b.w, 0.0f), h);
    boxes[i] = b;
}

__global__ void scale_flip_kernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t n, float sx,
                                  float sy, int flip, float net_w, float net_h) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = in[i];
    b.x *= sx; b.z *= sx; b.y *= sy; b.w *= sy;
    if (flip == 1) {
        const float x1 = net_w - b.z, x2 = net_w - b.x;
        b.x = x1; b.z = x2;
    } else if (flip == 2) {
        const float y1 = net_h - b.w, y2 = net_h - b.y;
        b.y = y1; b.w = y2;
    }
    out[i] = b;
}

// GDINO.resize_boxes (gdino.py:144-160) + Boxes.clip (gdino.py:135-136): normalised cxcywh -> pixel xyxy.
// Per box, in the reference's order: b *= (W, H, W, H); xy1 = c - wh / 2; xy2 = wh + xy1.
__global__ void cxcywh_to_xyxy_kernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t n, float W, float H,
                                      int clip) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = in[i];
    b.x *= W; b.y *= H; b.z *= W; b.w *= H;
    const float x1 = b.x - b.z / 2.0f, y1 = b.y - b.w / 2.0f;
    float4 o = make_float4(x1, y1, b.z + x1, b.w + y1);
    if (clip) {
        o.x = fminf(fmaxf(o.x, 0.0f), W); o.y = fminf(fmaxf(o.y, 0.0f), H);
        o.z = fminf(fmaxf(o.z, 0.0f), W); o.w = fminf(fmaxf(o.w, 0.0f), H);
    }
    out[i] = o;
}

}  // namespace coin
using namespace coin;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int coin_apply_deltas(const float* deltas, const float* boxes, float* out, int64_t R, int kreg,
                                 float wx, float wy, float ww, float wh, float scale_clamp, int clip,
                                 float clip_h, float clip_w, coin_stream_t stream) {
    COIN_REQUIRE(R >= 0 && kreg >= 1, "apply_deltas: bad sizes R=%lld kreg=%d", (long long)R, kreg);
    if (R == 0) return COIN_OK;
    COIN_REQUIRE(deltas && boxes && out, "apply_deltas: null pointer");
    COIN_REQUIRE(aligned16(deltas) && aligned16(boxes) && aligned16(out), "apply_deltas: pointers must be 16-byte aligned");
    const int64_t total = R * kreg;
    apply_deltas_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(deltas), reinterpret_cast<const float4*>(boxes),
        reinterpret_cast<float4*>(out), total, kreg, wx, wy, ww, wh, scale_clamp, clip, clip_h, clip_w);
    return check_launch("apply_deltas_kernel");
}

extern "C" int coin_get_deltas(const float* src, const float* tgt, float* out, int64_t F, float wx, float wy,
                               float ww, float wh, int32_t* invalid_flag, coin_stream_t stream) {
    COIN_REQUIRE(F >= 0, "get_deltas: bad size");
    if (F == 0) return COIN_OK;
    COIN_REQUIRE(src && tgt && out, "get_deltas: null pointer");
    COIN_REQUIRE(aligned16(src) && aligned16(tgt) && aligned16(out), "get_deltas: pointers must be 16-byte aligned");
    get_deltas_kernel<<<(unsigned)ceil_div(F, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(src), reinterpret_cast<const float4*>(tgt), reinterpret_cast<float4*>(out),
        F, wx, wy, ww, wh, invalid_flag);
    return check_launch("get_deltas_kernel");
}

extern "C" int coin_boxes_clip(float* boxes, int64_t n, float h, float w, coin_stream_t stream) {
    COIN_REQUIRE(n >= 0, "boxes_clip: bad size");
    if (n == 0) return COIN_OK;
    COIN_REQUIRE(boxes && aligned16(boxes), "boxes_clip: null or misaligned pointer");
    clip_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(boxes), n, h, w);
    return check_launch("clip_kernel");
}

extern "C" int coin_boxes_scale_flip(const float* in, float* out, int64_t n, float sx, float sy, int flip,
                                     float net_w, float net_h, coin_stream_t stream) {
    COIN_REQUIRE(n >= 0 && flip >= 0 && flip <= 2, "boxes_scale_flip: bad arguments");
    if (n == 0) return COIN_OK;
    COIN_REQUIRE(in && out && aligned16(in) && aligned16(out), "boxes_scale_flip: null or misaligned pointer");
    scale_flip_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n, sx, sy, flip, net_w, net_h);
    return check_launch("scale_flip_kernel");
}

extern "C" int coin_boxes_cxcywh_to_xyxy(const float* in, float* out, int64_t n, float img_h, float img_w, int clip,
                                         coin_stream_t stream) {
    COIN_REQUIRE(n >= 0, "boxes_cxcywh_to_xyxy: bad size");
    if (n == 0) return COIN_OK;
    COIN_REQUIRE(in && out && aligned16(in) && aligned16(out), "boxes_cxcywh_to_xyxy: null or misaligned pointer");
    cxcywh_to_xyxy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n, img_w, img_h, clip);
    return check_launch("cxcywh_to_xyxy_kernel");
}
