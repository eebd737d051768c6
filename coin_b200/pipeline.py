"""One step of COIN's RoI path over a batch of images, on one GPU.

This is the per-iteration work of ``CoinTrainer.run_step`` (coin/engine/trainer.py:160-218) that
lies between the backbone and the box-head GEMMs, in the reference's order:

  teacher branch (per image)
    T1  Box2BoxTransform.apply_deltas + Boxes.clip           fast_rcnn.py:729,145-147
    T2  score filter -> batched NMS -> top-100                fast_rcnn.py:116-175
    T3  cloud detections: Boxes.scale (+flip)                 base.py:80-136
    T4  knowledge separation, tags 'RCNN' and 'RPN'           trainer.py:338-478
  student branch (per image)
    S1  RPN proposal NMS (thr 0.7, keep[:post_nms_topk])      d2 find_top_rpn_proposals <- rpn.py:113
    S2  anchors vs A|C: pairwise_iou + Matcher(low quality)   rpn.py:209-228
    S3  proposals(+A,B) vs A|B|C: pairwise_iou + Matcher      clip_roi_heads.py:345-362
  student branch (batch)
    S4  ROIAlign forward on the sampled RoIs + on the C boxes clip_roi_heads.py:201-203,213-217
    S5  ROIAlign backward (gradient w.r.t. the res4 map)

Host round trips: two per BATCH (the detection counts after T2, the A/B/C counts after T4), each a
single small D2H copy; the reference has several per image (trainer.py:469, nonzero()/tolist()).
Everything else is asynchronous launches of libcoinops kernels on the current stream.
"""
import os
import ctypes
import struct
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, ops
from ._lib import check, lib
from .synth import Shape

ORIG_SCALE = 2048.0 / 1200.0  # Foggy-Cityscapes: 1024x2048 originals, 600x1200 network input


def anchors_for(shape: Shape) -> torch.Tensor:
    """detectron2 DefaultAnchorGenerator for the single stride-16 map (Base-Cloud.yaml:22-24):
    sizes 32..512 x ratios 0.5,1,2; (H, W, A) flattening. Host-side constant, built once."""
    import math
    base = []
    for s in (32, 64, 128, 256, 512):
        for r in (0.5, 1.0, 2.0):
            w = math.sqrt(s * s / r)
            h = r * w
            base.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    base = torch.tensor(base, dtype=torch.float32)
    hf, wf = shape.feat_hw
    sx = torch.arange(0, wf * shape.stride, step=shape.stride, dtype=torch.float32)
    sy = torch.arange(0, hf * shape.stride, step=shape.stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return (shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4)


def cloud_inputs_from_cache(cache, file_names, tag: str = "RCNN") -> Dict[str, torch.Tensor]:
    """The step's cloud-detection inputs (``{i}.cloud.gt_boxes / gt_classes / scores / probs``, boxes in the ORIGINAL image
    frame) taken from a device-resident ``coin_b200.cache.DetectionCache``: every tensor is a VIEW into the cache's flat
    arrays - no pickle, no deepcopy, no host-to-device copy (the reference: gdino_collector.py:86 deepcopy per lookup,
    trainer.py:457-459 ``.to(device)`` per step). ``update(d)`` the dict of ``RoIPathStep.h2d`` with it: the step's own T3
    stage (``process``, base.py:80-126) and knowledge separation then run on the cached rows directly."""
    out: Dict[str, torch.Tensor] = {}
    for i, name in enumerate(file_names):
        inst = cache.lookup(name, tag)
        out[f"{i}.cloud.gt_boxes"] = inst.pred_boxes.tensor
        out[f"{i}.cloud.gt_classes"] = inst.pred_classes
        out[f"{i}.cloud.scores"] = inst.scores
        out[f"{i}.cloud.probs"] = inst.probs
    return out


class RoIPathStep:
    BBOX_WEIGHTS = (10.0, 10.0, 5.0, 5.0)   # ROI_BOX_HEAD.BBOX_REG_WEIGHTS (fast_rcnn.py:297)
    SCORE_THRESH, NMS_THRESH, TOPK = 0.05, 0.5, 100
    RPN_NMS_THRESH = 0.7
    MATCH_THRESH = 0.5                      # CLOUD.MATCHER.IOU_THRESHOLDS (config.py:143)

    def __init__(self, shape: Shape, device, weight_for_box_a: float = 1.0, seed: int = 2024, share=None,
                 io_dtype: torch.dtype = torch.float32):
        """share: another RoIPathStep of the same shape whose constants (anchors, head gradient) are reused
        (a second graph instance for double buffering must not duplicate the 1.2 GB head gradient).
        io_dtype: dtype of the feature map coming in and of its gradient going out. The reference runs this path under
        autocast (trainer.py:175,187), where both are fp16; the arithmetic stays fp32 either way."""
        self.shape, self.device, self.w_a, self.io_dtype = shape, device, weight_for_box_a, io_dtype
        if share is not None:
            self.anchors, self.head_grad = share.anchors, share.head_grad
        else:
            self.anchors = anchors_for(shape).to(device)
            g = torch.Generator(device="cpu")
            g.manual_seed(seed + 1)
            # The box head behind ROIAlign is stood in for by a linear functional <pooled, G>: its gradient
            # w.r.t. the pooled features is the constant G, resident on the device like a weight.
            self.head_grad = torch.randn(shape.images * shape.rois, shape.channels, shape.pooled, shape.pooled,
                                         generator=g).to(device)
        self._streams: List[torch.cuda.Stream] = []
        self.roi_gate = os.environ.get("COIN_ROI_GATE", "none")   # none | det | det+rpn (see _run_static)
        # resident ROIAlign CTAs per SM inside the overlapped step (0: as many as fit = 4). With 4 the register file is
        # full and every short kernel of the other streams waits for a ROIAlign CTA to retire; 3 leaves them a quarter SM.
        self.roi_ctas_per_sm = int(os.environ.get("COIN_STEP_ROI_OCC", "0"))
        # preferred shared-memory carveout (%) of the ROIAlign kernels inside the step, -1: the driver's choice. The
        # driver sizes the carveout for the resident ROIAlign CTAs only, which keeps every kernel of the other streams
        # that needs more than ~26 KB of shared memory (sort, knowledge separation) off the SMs until ROIAlign drains.
        self.roi_carveout = int(os.environ.get("COIN_STEP_ROI_CARVEOUT", "-1"))
        self.c_mode = os.environ.get("COIN_STEP_C_MODE", "with_bwd")        # side | between | with_bwd (see _run_static)
        # launch order of the two big ROIAlign grids: the smallest p % of the RoIs go last (ops.roi_launch_order), 0: off
        self.roi_tail_pct = int(os.environ.get("COIN_STEP_ROI_TAIL_PCT", "20"))
        # The private-box forward. Private boxes can span the whole map (a clipped, mis-regressed detection): a register-tile
        # CTA then walks 256 channels of a 37 x 75 map for up to ~1 ms, and the step time must not depend on which images a
        # rank draws. Boxes above the size thresholds (px^2, 600 feature cells at stride 16; or a side in px) go to the
        # separable kernel, the rest to the register-tile kernel (profiles/r02_step_schedule.md section 3: 0.917-0.933 ms for
        # the data of all 8 ranks; 0.90 typical / 2.43 worst without the split). c_split_area = 0: no split, and launches of
        # fewer than c_reg_mink RoIs take the separable kernel entirely (round 1's choice: 0.94-0.97 ms).
        self.c_split_area = float(os.environ.get("COIN_STEP_C_SPLIT_AREA", str(600 * 256)))
        self.c_split_side = float(os.environ.get("COIN_STEP_C_SPLIT_SIDE", "640"))
        self.c_reg_mink = int(os.environ.get("COIN_STEP_C_REG_MINK", "1024"))
        self.c_chans = int(os.environ.get("COIN_STEP_C_CHANS", "256"))          # channels per register-tile CTA of that launch
        self.overlap = True         # issue independent stages on side streams (False: everything on the caller's stream)
        self.timeline = None        # tools/step_timeline.py: dict name -> external CUDA event recorded in the graph
        self.kernel_events = None   # bench.py: {"fwd": [], "bwd": []} to time the two dominant kernels live

    # -- data movement -------------------------------------------------------------------------
    def host_inputs(self, batch) -> Dict[str, torch.Tensor]:
        """Flat dict of the per-step INPUT tensors (pinned host memory) for the end-to-end timing."""
        flat = {"features": batch["features"].to(self.io_dtype)}
        for i, img in enumerate(batch["images"]):
            for k in ("teacher_rois", "teacher_deltas", "teacher_probs", "proposals", "rois", "rpn_boxes",
                      "rpn_scores"):
                flat[f"{i}.{k}"] = img[k]
            for side in ("cloud", "clip"):
                if side == "clip":
                    continue  # CLIP-detector detections are produced on the device by T2
                for k, v in img[side].items():
                    flat[f"{i}.{side}.{k}"] = v * ORIG_SCALE if k == "gt_boxes" else v
        return {k: v.contiguous().pin_memory() for k, v in flat.items()}

    def h2d(self, pinned: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {k: v.to(self.device, non_blocking=True) for k, v in pinned.items()}

    def to_device(self, batch):
        return self.h2d(self.host_inputs(batch))

    @staticmethod
    def input_bytes(pinned) -> int:
        return sum(v.numel() * v.element_size() for v in pinned.values())

    # -- the step --------------------------------------------------------------------------------
    def _side_streams(self, n: int) -> List[torch.cuda.Stream]:
        """Stream 0 carries the HBM-bound ROIAlign forward/backward at normal priority; the others carry the
        short latency-bound kernels at HIGH priority, so their CTAs are scheduled as soon as a ROIAlign CTA
        retires instead of queueing behind the whole ROIAlign grid."""
        while len(self._streams) < n:
            prio = 0 if not self._streams else -1
            self._streams.append(torch.cuda.Stream(device=self.device, priority=prio))
        return self._streams[:n]

    def run(self, d: Dict[str, torch.Tensor], backward: bool = True) -> Dict[str, object]:
        """One step. The stages of different images (and the two NMS chains of one image) are independent,
        so they are issued on separate CUDA streams: the latency-bound single-CTA kernels of one image
        overlap those of the others, and the HBM-bound ROIAlign forward/backward (which depend only on the
        feature map and the sampled RoIs) overlap the whole teacher/matching branch. Every side stream is
        joined into the caller's stream before run() returns."""
        sh, dev = self.shape, self.device
        n_img = sh.images
        img_size = (sh.height, sh.width)
        main = torch.cuda.current_stream()
        if not self.overlap:
            return self._run(d, backward, [main] * (3 * n_img + 1))
        side = self._side_streams(3 * n_img + 1)
        start = main.record_event()
        for st in side:
            st.wait_event(start)
        out = self._run(d, backward, side)
        for st in side:
            main.wait_stream(st)
        return out

    def _run(self, d, backward, streams) -> Dict[str, object]:
        sh, dev = self.shape, self.device
        n_img = sh.images
        img_size = (sh.height, sh.width)
        main = torch.cuda.current_stream()
        s_roi, s_img = streams[0], streams[1:]
        scale = (1.0 / sh.stride,)
        size = (sh.pooled, sh.pooled)
        ev = self.kernel_events
        out: Dict[str, object] = {"dets": [], "abc": [], "roi_labels": [], "rpn_labels": [], "rpn_keep": []}

        # ---- ROIAlign forward (S4) and backward (S5) over the sampled RoIs: own stream, needs only the map
        with torch.cuda.stream(s_roi):
            nhwc = ops.to_nhwc_f32(d["features"])
            rois = self._batched_rois(d, n_img)
            perm, plan = ops.roi_launch_plan(rois, scale[0], small_pct=self.roi_tail_pct)
            out["pooled"] = ops.roi_align_forward_planned([nhwc], scale, rois, size, 0, True, torch.float32, plan=plan,
                                                          events=ev["fwd"] if ev else None)
            if backward:
                n, c, h, w = d["features"].shape
                out["grad_features"] = ops.roi_align_backward(self.head_grad, [(n, c, h, w)], scale, rois, None, size,
                                                              0, True, [self.io_dtype],
                                                              events=ev["bwd"] if ev else None, perm=perm)[0]

        # ---- teacher branch and RPN NMS: one stream per image and chain, no host sync inside the loop
        dets, cloud_boxes, rpn = [None] * n_img, [None] * n_img, [None] * n_img
        for i in range(n_img):
            with torch.cuda.stream(s_img[i]):
                dec = ops.apply_deltas(d[f"{i}.teacher_deltas"], d[f"{i}.teacher_rois"], self.BBOX_WEIGHTS,
                                       clip_to=img_size)                                                    # T1
                dets[i] = ops.det_postprocess(dec, d[f"{i}.teacher_probs"], img_size, self.SCORE_THRESH,
                                              self.NMS_THRESH, self.TOPK, sync=False)                       # T2
                cloud_boxes[i] = ops.boxes_scale_flip(d[f"{i}.cloud.gt_boxes"], sh.width / (sh.width * ORIG_SCALE),
                                                      sh.height / (sh.height * ORIG_SCALE), "no", img_size)  # T3
            with torch.cuda.stream(s_img[n_img + i]):
                rpn[i] = ops.batched_nms(d[f"{i}.rpn_boxes"], d[f"{i}.rpn_scores"], None, self.RPN_NMS_THRESH,
                                         "plain", sh.rpn_post_nms, sync=False)                              # S1
        for st in s_img[: 2 * n_img]:
            main.wait_stream(st)
        counts1 = torch.stack([x[5] for x in dets] + [x[1] for x in rpn]).view(-1).cpu()                 # sync 1
        n_det = counts1[:n_img].tolist()
        n_rpn = counts1[n_img:].tolist()

        # ---- knowledge separation (T4): one launch per image and tag, each on its own stream
        raws = []
        for i in range(n_img):
            b, s, p, c, _, _ = dets[i]
            nd = n_det[i]
            for t, tag in enumerate(("RCNN", "RPN")):
                st = s_img[2 * i + t]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    raws.append(ops.match_abc(cloud_boxes[i], d[f"{i}.cloud.gt_classes"], d[f"{i}.cloud.scores"],
                                              b[:nd], c[:nd], s[:nd], tag, self.MATCH_THRESH, self.w_a, sync=False))
        for st in s_img[: 2 * n_img]:
            main.wait_stream(st)
        counts2 = torch.stack([r["counts"] for r in raws]).cpu().tolist()                                 # sync 2

        c_rois = [None] * n_img
        for i in range(n_img):
            st = s_img[2 * n_img + i]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                b, s, p, c, roi_idx, _ = dets[i]
                nd = n_det[i]
                out["dets"].append({"pred_boxes": b[:nd], "scores": s[:nd], "probs": p[:nd], "pred_classes": c[:nd],
                                    "roi_index": roi_idx[:nd]})
                out["rpn_keep"].append(rpn[i][0][: n_rpn[i]])
                cloud = {"gt_boxes": cloud_boxes[i], "gt_classes": d[f"{i}.cloud.gt_classes"],
                         "scores": d[f"{i}.cloud.scores"], "probs": d[f"{i}.cloud.probs"]}
                clip = {"gt_boxes": b[:nd], "gt_classes": c[:nd], "scores": s[:nd], "probs": p[:nd]}
                per_tag = {}
                for t, tag in enumerate(("RCNN", "RPN")):
                    r = ops.match_abc_narrow(raws[2 * i + t], counts2[2 * i + t])
                    per_tag[tag] = self._pack(r, cloud, clip, tag)
                out["abc"].append(per_tag)

                a, bb, cc = per_tag["RCNN"]
                gt = torch.cat((a["gt_boxes"], bb["gt_boxes"], cc["gt_boxes"]))
                props = torch.cat((d[f"{i}.proposals"], a["gt_boxes"], bb["gt_boxes"]))   # add_ground_truth_to_proposals
                idx, lab = ops.iou_match(gt, props, [0.5], [0, 1], False)                                  # S3
                la, lb, lc = a["gt_boxes"].shape[0], bb["gt_boxes"].shape[0], cc["gt_boxes"].shape[0]
                ops.relabel_roi_(idx, lab, la + lb, la + lb + lc)
                out["roi_labels"].append((idx, lab))

                a2, _, c2 = per_tag["RPN"]
                gt2 = torch.cat((a2["gt_boxes"], c2["gt_boxes"]))
                idx2, lab2 = ops.iou_match(gt2, self.anchors, [0.3, 0.7], [0, -1, 1], True)                # S2
                out["rpn_labels"].append(ops.relabel_rpn_(idx2, lab2, a2["gt_boxes"].shape[0], c2["gt_boxes"].shape[0]))
                cb = cc["gt_boxes"]
                c_rois[i] = torch.cat((torch.full((cb.shape[0], 1), float(i), device=dev), cb), dim=1)

        # ---- ROIAlign forward on the private (C) boxes, behind the big forward/backward on the RoI stream
        for st in s_img[2 * n_img:]:
            s_roi.wait_stream(st)
        with torch.cuda.stream(s_roi):
            out["pooled_c"] = ops.roi_align_forward([nhwc], scale, torch.cat(c_rois), None, size, 0, True, torch.float32)
        out["summary"] = {"dets": n_det, "rpn_keep": n_rpn, "abc": [c2[:3] for c2 in counts2]}
        return out

    # -- sync-free step: fixed launch sequence, device-side lengths, CUDA-graph capturable ------------
    def run_static(self, d: Dict[str, torch.Tensor], backward: bool = True) -> Dict[str, object]:
        """The same step as run() with NO host round trip: every variable-length set lives in a worst-case
        buffer with its length in device memory (the *_dev entry points of libcoinops), so the launch
        sequence is independent of the data and can be captured in a CUDA graph (capture()/replay()).
        Returns padded tensors plus ``counts`` (one device int32 vector); finalize() reads the counts back
        once and narrows the buffers to the dict format of run()."""
        main = torch.cuda.current_stream()
        n_img = self.shape.images
        side = self._side_streams(3 * n_img + 1) if self.overlap else [main] * (3 * n_img + 1)
        start = main.record_event()
        for st in side:
            if st is not main:
                st.wait_event(start)
        out = self._run_static(d, backward, side)
        for st in side:
            if st is not main:
                main.wait_stream(st)
        return out

    @staticmethod
    def _batched_rois(d, n_img: int) -> torch.Tensor:
        """[batch index | box] rows of every image's sampled RoIs (convert_boxes_to_pooler_format) in ONE launch: the ROIAlign
        forward waits for this list and its launch plan, so the seven torch fill / cat kernels it used to take were ~25 us
        of lead-in at the start of every step."""
        return ops.concat_rows([(d[f"{i}.rois"], None, float(i)) for i in range(n_img)], width_out=5)[0]

    def _mark(self, name: str) -> None:
        """Named timestamp on the current stream (only when a timeline is being collected)."""
        if self.timeline is not None:
            e = torch.cuda.Event(enable_timing=True, external=torch.cuda.is_current_stream_capturing())
            e.record()
            self.timeline[name] = e

    def _run_static(self, d, backward, streams) -> Dict[str, object]:
        sh, dev = self.shape, self.device
        n_img = sh.images
        img_size = (sh.height, sh.width)
        s_roi, s_img = streams[0], streams[1:]
        scale = (1.0 / sh.stride,)
        size = (sh.pooled, sh.pooled)
        out: Dict[str, object] = {"dets": [], "abc": [], "roi_labels": [], "rpn_labels": [], "rpn_keep": []}
        counts: List[torch.Tensor] = []     # device int32 tensors, concatenated at the end
        slots: Dict[str, int] = {}          # name -> offset into the concatenated counts vector

        def slot(name, t):
            slots[name] = sum(int(c.numel()) for c in counts)
            counts.append(t.view(-1))

        self._mark("start")
        with torch.cuda.stream(s_img[2 * n_img]):     # (this stream's chain starts only after knowledge separation)
            rois = self._batched_rois(d, n_img)
            # launch order (smallest RoIs last) + size split (map-sized RoIs, if any, go to the separable kernel:
            # ops.roi_align_forward_planned) in one launch, beside the layout transform below
            perm, plan = ops.roi_launch_plan(rois, scale[0], small_pct=self.roi_tail_pct)
            rois_ready = s_img[2 * n_img].record_event()
            gbuf = gbuf_ready = None
            if backward:      # the backward's accumulation buffer is zeroed here, beside the forward, not between the two
                n, c, h, w = d["features"].shape
                gbuf = torch.zeros((n, h, w, c), dtype=torch.float32, device=dev)
                gbuf_ready = s_img[2 * n_img].record_event()
        with torch.cuda.stream(s_roi):
            nhwc = ops.to_nhwc_f32(d["features"])
            nhwc_ready = s_roi.record_event()
            s_roi.wait_event(rois_ready)

        # ---- teacher detections (T1-T3) and RPN NMS (S1): one stream per image and chain
        clouds, clips, ndets, det_done, gate = [], [], [], [], []
        for i in range(n_img):
            st_t, st_p = s_img[i], s_img[n_img + i]
            with torch.cuda.stream(st_p):                                                                 # S1
                keep, nkeep = ops.batched_nms(d[f"{i}.rpn_boxes"], d[f"{i}.rpn_scores"], None, self.RPN_NMS_THRESH,
                                              "plain", sh.rpn_post_nms, sync=False)
                self._mark(f"img{i}.rpn_nms_done")
                if self.roi_gate == "det+rpn":
                    gate.append(st_p.record_event())
            with torch.cuda.stream(st_t):
                dec = ops.apply_deltas(d[f"{i}.teacher_deltas"], d[f"{i}.teacher_rois"], self.BBOX_WEIGHTS,
                                       clip_to=img_size)                                                    # T1
                b, s, p, c, roi_idx, ndet = ops.det_postprocess(dec, d[f"{i}.teacher_probs"], img_size,
                                                                self.SCORE_THRESH, self.NMS_THRESH, self.TOPK,
                                                                sync=False)                                 # T2
                cloud = {"gt_boxes": ops.boxes_scale_flip(d[f"{i}.cloud.gt_boxes"], sh.width / (sh.width * ORIG_SCALE),
                                                          sh.height / (sh.height * ORIG_SCALE), "no", img_size),  # T3
                         "gt_classes": d[f"{i}.cloud.gt_classes"], "scores": d[f"{i}.cloud.scores"],
                         "probs": d[f"{i}.cloud.probs"]}
                self._mark(f"img{i}.det_done")
                det_done.append(st_t.record_event())
            clouds.append(cloud)
            clips.append({"gt_boxes": b, "gt_classes": c, "scores": s, "probs": p})
            ndets.append(ndet)
            out["dets"].append({"pred_boxes": b, "scores": s, "probs": p, "pred_classes": c, "roi_index": roi_idx})
            out["rpn_keep"].append(keep)
            slot(f"det{i}", ndet)
            slot(f"rpn{i}", nkeep)
        if self.roi_gate in ("det", "det+rpn"):
            gate += det_done
        out["_keepalive"] = (clouds, clips, gbuf)   # read by other streams: must outlive this function's locals

        # ---- ROIAlign forward (S4) and backward (S5) over the sampled RoIs: needs only the map. roi_gate
        #      optionally holds it back until the short, resource-hungry kernels of the chains above (sorts,
        #      masks) have had the machine to themselves.
        for e in gate:
            s_roi.wait_event(e)
        occ = self.roi_ctas_per_sm if self.overlap else 0
        roi_opts = dict(COIN_ROI_CTAS_PER_SM=occ, COIN_ROI_CARVEOUT=self.roi_carveout if self.overlap else -1)
        with torch.cuda.stream(s_roi), _lib.options(**roi_opts):
            self._mark("roi.begin_fwd")
            ev = self.kernel_events
            out["pooled"] = ops.roi_align_forward_planned([nhwc], scale, rois, size, 0, True, torch.float32, plan=plan,
                                                          events=ev["fwd"] if ev else None,
                                                          big_stream=s_img[2 * n_img] if self.overlap else None)
            self._mark("roi.end_fwd")
            fwd_done = s_roi.record_event()

        def run_backward():
            s_roi.wait_event(gbuf_ready)
            with torch.cuda.stream(s_roi), _lib.options(**roi_opts):
                n, c, h, w = d["features"].shape
                out["grad_features"] = ops.roi_align_backward(self.head_grad, [(n, c, h, w)], scale, rois, None, size,
                                                              0, True, [self.io_dtype],
                                                              events=ev["bwd"] if ev else None, perm=perm, zeroed=[gbuf])[0]
                self._mark("roi.end_bwd")
        if backward and self.c_mode != "between":
            run_backward()

        # ---- knowledge separation (T4) and labelling (S3, S2): tag RCNN continues on the teacher stream,
        #      tag RPN on the image's third stream
        c_segs = []
        for i in range(n_img):
            cloud, clip, ndet = clouds[i], clips[i], ndets[i]
            per_tag = {}
            st_t, st_r = s_img[i], s_img[2 * n_img + i]
            with torch.cuda.stream(st_t):
                # T4: ONE launch separates the knowledge for both tags (they share everything up to the A/B split)
                both = ops.match_abc_fields_both_dev(cloud, clip, ndet, self.MATCH_THRESH, self.w_a)
                self._mark(f"img{i}.abc_done")
                abc_done = st_t.record_event()
            for tag, st in (("RCNN", st_t), ("RPN", st_r)):
                st.wait_event(abc_done)
                with torch.cuda.stream(st):
                    a, bb, cc, cnt = both[tag]
                    slot(f"abc{i}.{tag}", cnt)
                    per_tag[tag] = (a, bb, cc)
                    n_a, n_b, n_c = cnt[0:1], cnt[1:2], cnt[2:3]
                    if tag == "RCNN":
                        # the private-box ROIAlign only needs the C set: it may start before the labelling below
                        c_segs.append((cc["gt_boxes"], n_c, float(i), st.record_event()))
                        gt, n_gt = ops.concat_rows([(a["gt_boxes"], n_a, 0.0), (bb["gt_boxes"], n_b, 0.0),
                                                    (cc["gt_boxes"], n_c, 0.0)])
                        props, n_props = ops.concat_rows([(d[f"{i}.proposals"], None, 0.0), (a["gt_boxes"], n_a, 0.0),
                                                          (bb["gt_boxes"], n_b, 0.0)])   # add_ground_truth_to_proposals
                        idx, lab = ops.iou_match_dev(gt, n_gt, props, n_props, [0.5], [0, 1], False)       # S3
                        ops.relabel_roi_dev_(idx, lab, n_props, n_a, n_b, n_c)
                        out["roi_labels"].append((idx, lab))
                        slot(f"props{i}", n_props)
                        self._mark(f"img{i}.roi_labels_done")
                    else:
                        gt2, n_gt2 = ops.concat_rows([(a["gt_boxes"], n_a, 0.0), (cc["gt_boxes"], n_c, 0.0)])
                        idx2, lab2 = ops.iou_match_dev(gt2, n_gt2, self.anchors, None, [0.3, 0.7], [0, -1, 1], True)  # S2
                        out["rpn_labels"].append(ops.relabel_rpn_dev_(idx2, lab2, n_a, n_c))
                        self._mark(f"img{i}.rpn_labels_done")
            out["abc"].append(per_tag)
            out.setdefault("_keepalive2", []).append(both)

        # ---- ROIAlign forward on the private (C) boxes of every image (known long before the big forward ends).
        #      c_mode "side": on a high-priority stream as soon as the boxes exist (competes with the forward);
        #      "between": on the ROIAlign stream between the forward and the backward; "with_bwd" (default, measured best:
        #      profiles/r02_step_schedule.md): high-priority stream, released when the forward is through, so that it
        #      shares the machine with the backward - two latency-bound kernels fill each other's gaps
        s_c = s_roi if self.c_mode == "between" else s_img[0]
        for seg in c_segs:
            s_c.wait_event(seg[3])
        s_c.wait_event(nhwc_ready)
        if self.c_mode == "with_bwd":      # high-priority stream, but only once the forward is through: shares the
            s_c.wait_event(fwd_done)       # machine with the backward (two latency-bound kernels fill each other's gaps)
        with torch.cuda.stream(s_c):
            c_rois, n_c_rois = ops.concat_rows([seg[:3] for seg in c_segs], width_out=5)
            if self.c_split_area > 0:
                # private boxes can span the whole map (a clipped, mis-regressed detection); one register-tile CTA then
                # walks 256 channels of a 37 x 75 map for ~1 ms. The few boxes above the size thresholds are pooled by the
                # separable kernel (first: they take longest), everything else by the register-tile kernel, same output.
                plan_c = ops.roi_split_by_area(c_rois, n_c_rois, self.c_split_area, self.c_split_side, None, ops.BIG_ROI_CAP)
                with _lib.options(COIN_ROI_REG_CHANS_SMALL=self.c_chans):
                    out["pooled_c"] = ops.roi_align_forward_planned([nhwc], scale, c_rois, size, 0, True, torch.float32,
                                                                    plan=plan_c,
                                                                    big_stream=s_img[2 * n_img + 1] if self.overlap and n_img > 1 else None)
            else:
                with _lib.options(COIN_ROI_REG_MINK=self.c_reg_mink, COIN_ROI_REG_CHANS_SMALL=self.c_chans):
                    out["pooled_c"] = ops.roi_align_forward([nhwc], scale, c_rois, None, size, 0, True, torch.float32,
                                                            k_dev=n_c_rois)
            slot("c_rois", n_c_rois)
            self._mark("roi.pooled_c_done")
            cat_done = s_c.record_event()
        if backward and self.c_mode == "between":
            run_backward()
        torch.cuda.current_stream().wait_event(cat_done)
        for st in streams:
            if st is not torch.cuda.current_stream():
                torch.cuda.current_stream().wait_stream(st)
        out["counts"] = torch.cat(counts)
        self._mark("end")
        out["slots"] = slots
        return out

    def finalize(self, out: Dict[str, object], counts_host=None) -> Dict[str, object]:
        """One D2H read of the counts vector, then narrow every padded buffer: the result has the format
        of run() (and of oracle/pipeline_ref.run)."""
        cnt = (out["counts"].cpu() if counts_host is None else counts_host).tolist()
        sl = out["slots"]
        n_img = self.shape.images
        res: Dict[str, object] = {"dets": [], "abc": [], "roi_labels": [], "rpn_labels": out["rpn_labels"],
                                  "rpn_keep": [], "pooled": out["pooled"]}
        if "grad_features" in out:
            res["grad_features"] = out["grad_features"]
        summary_abc = []
        for i in range(n_img):
            nd = cnt[sl[f"det{i}"]]
            res["dets"].append({k: v[:nd] for k, v in out["dets"][i].items()})
            res["rpn_keep"].append(out["rpn_keep"][i][: cnt[sl[f"rpn{i}"]]])
            per_tag = {}
            for tag in ("RCNN", "RPN"):
                o = sl[f"abc{i}.{tag}"]
                na, nb, ncc, status = cnt[o], cnt[o + 1], cnt[o + 2], cnt[o + 3]
                if status & 16:
                    raise AssertionError("match_abc: a cloud self-cluster has a single class (util.py:488 assert)")
                if status & 8:
                    raise AssertionError("match_abc: a duplicate group holds several boxes of the matched class")
                if status & 4:
                    raise RuntimeError("match_abc: pair capacity exceeded")
                a, b, c = out["abc"][i][tag]
                per_tag[tag] = ({k: v[:na] for k, v in a.items()},
                                None if b is None else {k: v[:nb] for k, v in b.items()},
                                {k: v[:ncc] for k, v in c.items()})
                summary_abc.append([na, nb, ncc])
            res["abc"].append(per_tag)
            m = cnt[sl[f"props{i}"]]
            idx, lab = out["roi_labels"][i]
            res["roi_labels"].append((idx[:m], lab[:m]))
        res["pooled_c"] = out["pooled_c"][: cnt[sl["c_rois"]]]
        res["summary"] = {"dets": [cnt[sl[f"det{i}"]] for i in range(n_img)],
                          "rpn_keep": [cnt[sl[f"rpn{i}"]] for i in range(n_img)], "abc": summary_abc}
        return res

    def time_roi_kernels(self, d, events: Dict[str, list], iters: int = 10, warmup: int = 3) -> None:
        """Launches the step's two dominant kernels (ROIAlign forward / backward over the sampled RoIs) alone
        on the current stream, `iters` times, appending a CUDA-event pair per launch to events['fwd'/'bwd']."""
        sh, dev = self.shape, self.device
        nhwc = ops.to_nhwc_f32(d["features"])
        rois = self._batched_rois(d, sh.images)
        n, c, h, w = d["features"].shape
        size, scale = (sh.pooled, sh.pooled), (1.0 / sh.stride,)
        perm = ops.roi_launch_order(rois, small_pct=self.roi_tail_pct)     # as in the step (its own small launch)
        for it in range(warmup + iters):
            ev = events if it >= warmup else {"fwd": None, "bwd": None}
            ops.roi_align_forward([nhwc], scale, rois, None, size, 0, True, torch.float32, events=ev["fwd"], perm=perm)   # (the
            # step's split finds no map-sized RoI among the sampled ones: this IS the launch it makes)
            ops.roi_align_backward(self.head_grad, [(n, c, h, w)], scale, rois, None, size, 0, True, [torch.float32],
                                   events=ev["bwd"], perm=perm)

    # -- CUDA graph ---------------------------------------------------------------------------------
    def capture(self, d: Dict[str, torch.Tensor], backward: bool = True, warmup: int = 2, keep_graph: bool = False):
        """Captures run_static over the (static) input tensors ``d`` into a CUDA graph. Later steps copy
        new inputs into the same tensors (copy_inputs) and call replay()."""
        self._graph_in = d
        cap_stream = torch.cuda.Stream(device=self.device)
        cap_stream.wait_stream(torch.cuda.current_stream())
        ev, self.kernel_events = self.kernel_events, None   # timing events belong to the captured launches only
        with torch.cuda.stream(cap_stream):
            for _ in range(warmup):       # allocator warm-up and lazy module loading outside the capture
                self.run_static(d, backward)
        self.kernel_events = ev
        torch.cuda.current_stream().wait_stream(cap_stream)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph(keep_graph=keep_graph) if keep_graph else torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=cap_stream):
            self._graph_out = self.run_static(d, backward)
        return self._graph_out

    def copy_inputs(self, src: Dict[str, torch.Tensor]) -> None:
        """Copies one step's inputs (pinned host or device tensors) into the captured graph's input tensors."""
        for k, v in self._graph_in.items():
            v.copy_(src[k], non_blocking=True)

    def replay(self) -> Dict[str, object]:
        self._graph.replay()
        return self._graph_out

    @staticmethod
    def _pack(r, cloud, clip, tag):
        """Gather the reference's A / B / C fields (trainer.py:393-455) from the index lists."""
        def side(on_idx, off_idx, boxes, split):
            o = {"gt_boxes": boxes}
            if split:
                o["gt_classes_offline"], o["gt_classes_online"] = clip["gt_classes"][off_idx], cloud["gt_classes"][on_idx]
            else:
                o["gt_classes"] = clip["gt_classes"][off_idx]
            o["gt_scores_online"], o["gt_scores_offline"] = cloud["scores"][on_idx], clip["scores"][off_idx]
            o["gt_probs_online"], o["gt_probs_offline"] = cloud["probs"][on_idx], clip["probs"][off_idx]
            return o

        a = side(r["a_on"], r["a_off"], r["a_boxes"], False)
        b = side(r["b_on"], r["b_off"], r["b_boxes"], True) if tag == "RCNN" else None
        off_rows, on_rows = r["c_off"], r["c_on"]
        c = {"gt_boxes": torch.cat((clip["gt_boxes"][off_rows], cloud["gt_boxes"][on_rows])),
             "gt_classes": torch.cat((clip["gt_classes"][off_rows], cloud["gt_classes"][on_rows])),
             "gt_scores": torch.cat((clip["scores"][off_rows], cloud["scores"][on_rows])),
             "gt_probs": torch.cat((clip["probs"][off_rows], cloud["probs"][on_rows]))}
        return a, b, c

    # -- results that travel back to the host in the end-to-end measurement ----------------------
    @staticmethod
    def result_spec(out) -> List[Tuple[torch.Tensor, Optional[int]]]:
        """The un-narrowed result buffers of a sync-free step in the order of ``result_tensors(finalize(out))`` (without the
        feature-map gradient), each with the index of its live row count in ``out['counts']`` (None: every row is live)."""
        sl = out["slots"]
        spec: List[Tuple[torch.Tensor, Optional[int]]] = []
        for i, dd in enumerate(out["dets"]):
            spec += [(v, sl[f"det{i}"]) for v in dd.values()]
        for i, per_tag in enumerate(out["abc"]):
            for tag in ("RCNN", "RPN"):
                o = sl[f"abc{i}.{tag}"]
                for part, ci in zip(per_tag[tag], (o, o + 1, o + 2)):
                    if part is not None:
                        spec += [(v, ci) for v in part.values()]
        for i, (idx, lab) in enumerate(out["roi_labels"]):
            spec += [(idx, sl[f"props{i}"]), (lab, sl[f"props{i}"])]
        for tup in out["rpn_labels"]:
            spec += [(t, None) for t in tup]
        spec += [(k, sl[f"rpn{i}"]) for i, k in enumerate(out["rpn_keep"])]
        return spec

    @staticmethod
    def result_tensors(out) -> List[torch.Tensor]:
        res = []
        for dd in out["dets"]:
            res += list(dd.values())
        for per_tag in out["abc"]:
            for tag in ("RCNN", "RPN"):
                for part in per_tag[tag]:
                    if part is not None:
                        res += list(part.values())
        for idx, lab in out["roi_labels"]:
            res += [idx, lab]
        for tup in out["rpn_labels"]:
            res += list(tup)
        res += out["rpn_keep"]
        if "grad_features" in out:
            res.append(out["grad_features"])
        return res


class PipelinedSteps:
    """End-to-end execution of consecutive steps with double buffering: while the graph of step n runs, the
    inputs of step n+1 travel host -> device and the results of step n-1 travel device -> host (four
    streams, two graph instances). Every step still pays its own H2D copy from pinned host memory, its own
    length read-back and its own D2H copy of the live results; only their latency is overlapped."""

    def __init__(self, first: RoIPathStep, d_first: Dict[str, torch.Tensor], backward: bool = True):
        dev = first.device
        self.slots = [first, RoIPathStep(first.shape, dev, first.w_a, share=first, io_dtype=first.io_dtype)]
        # One contiguous device buffer per dtype and slot holds all inputs of a step (the graphs' input tensors
        # are views into it), mirrored by one pinned host buffer per dtype: a step's inputs travel in one H2D
        # copy per dtype instead of ~40 small ones.
        layout: Dict[torch.dtype, list] = {}
        for k, v in d_first.items():
            layout.setdefault(v.dtype, []).append((k, tuple(v.shape), v.numel()))
        self.host_packed = {dt: torch.empty((sum(n for _, _, n in items),), dtype=dt).pin_memory()
                            for dt, items in layout.items()}
        self.host_in: Dict[str, torch.Tensor] = {}
        self.dev_packed = []
        dev_in = []
        for s in range(2):
            bufs = {dt: torch.empty((sum(n for _, _, n in items),), dtype=dt, device=dev) for dt, items in layout.items()}
            views = {}
            for dt, items in layout.items():
                off = 0
                for k, shape, n in items:
                    views[k] = bufs[dt][off: off + n].view(shape)
                    if s == 0:
                        self.host_in[k] = self.host_packed[dt][off: off + n].view(shape)
                    off += n
            for k, v in d_first.items():
                views[k].copy_(v)
            self.dev_packed.append(bufs)
            dev_in.append(views)
        for slot, views in zip(self.slots, dev_in):
            slot.capture(views, backward)
        self.s_in, self.s_c, self.s_out, self.s_p = (torch.cuda.Stream(device=dev) for _ in range(4))
        self.counts_host = [torch.empty(s._graph_out["counts"].shape, dtype=torch.int32).pin_memory() for s in self.slots]
        # The ~100 variable-length results of a step (detections, A/B/C fields, labels, keep lists) are packed ON THE DEVICE into
        # one staging buffer (coin_pack_rows: live prefixes only, lengths read on the device) and leave in ONE copy; the host
        # slices the pinned copy lazily. (Narrowing and packing them in Python cost 0.6 ms per step - more than the link - and
        # the slot's next graph had to wait for it.)
        self.specs, self.items_dev, self.packed_dev, self.packed_host, self.offsets_dev, self.offsets_host = [], [], [], [], [], []
        for slot in self.slots:
            spec = slot.result_spec(slot._graph_out)
            table, cap = b"", 0
            for t, ci in spec:
                assert t.is_contiguous()
                rows = int(t.shape[0]) if t.dim() else 1
                row_bytes = (t.numel() // max(rows, 1)) * t.element_size()
                table += struct.pack("<QqqiI", t.data_ptr(), row_bytes, rows, -1 if ci is None else int(ci), 0)
                cap += (rows * row_bytes + 15) // 16 * 16
            self.specs.append(spec)
            self.items_dev.append(torch.frombuffer(bytearray(table), dtype=torch.uint8).to(dev))
            self.packed_dev.append(torch.empty((max(cap, 16),), dtype=torch.uint8, device=dev))
            self.packed_host.append(torch.empty((max(cap, 16),), dtype=torch.uint8).pin_memory())
            self.offsets_dev.append(torch.zeros((len(spec) + 1,), dtype=torch.int64, device=dev))
            self.offsets_host.append(torch.zeros((len(spec) + 1,), dtype=torch.int64).pin_memory())
        self.grad_host = [None, None]
        self.h2d_done = [None, None]
        self.compute_done = [None, None]
        self.d2h_done = [None, None]
        self.d2h_bytes = 0

    def load_inputs(self, src: Dict[str, torch.Tensor]) -> None:
        """Host-side: writes one step's inputs into the pinned staging buffers (what a data loader would fill)."""
        for k, v in self.host_in.items():
            v.copy_(src[k])

    def _h2d(self, n, pinned):
        s = n % 2
        if pinned is not None:
            self.load_inputs(pinned)
        with torch.cuda.stream(self.s_in):
            if self.compute_done[s] is not None:
                self.s_in.wait_event(self.compute_done[s])      # the graph of step n-2 has read these inputs
            for dt, buf in self.dev_packed[s].items():
                buf.copy_(self.host_packed[dt], non_blocking=True)
            self.h2d_done[s] = self.s_in.record_event()

    def _compute(self, n):
        s = n % 2
        with torch.cuda.stream(self.s_c):
            self.s_c.wait_event(self.h2d_done[s])
            if self.d2h_done[s] is not None:
                self.s_c.wait_event(self.d2h_done[s])           # the results of step n-2 have left these buffers
            out = self.slots[s].replay()
            graph_done = self.s_c.record_event()
        with torch.cuda.stream(self.s_p):      # packing and the two tiny read-backs: beside the next step's graph, not before it
            self.s_p.wait_event(graph_done)
            check(lib.coin_pack_rows(ctypes.c_void_p(self.items_dev[s].data_ptr()), len(self.specs[s]),
                                     ctypes.c_void_p(out["counts"].data_ptr()), ctypes.c_void_p(self.packed_dev[s].data_ptr()),
                                     self.packed_dev[s].numel(), ctypes.c_void_p(self.offsets_dev[s].data_ptr()),
                                     ctypes.c_void_p(self.s_p.cuda_stream)))
            self.counts_host[s].copy_(out["counts"], non_blocking=True)
            self.offsets_host[s].copy_(self.offsets_dev[s], non_blocking=True)
            self.compute_done[s] = self.s_p.record_event()

    def _d2h(self, n):
        s = n % 2
        self.compute_done[s].synchronize()                      # the lengths of step n are on the host
        total = int(self.offsets_host[s][-1])
        nbytes = self.counts_host[s].numel() * 4 + self.offsets_host[s].numel() * 8 + total
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.compute_done[s])
            self.packed_host[s][:total].copy_(self.packed_dev[s][:total], non_blocking=True)
            g = self.slots[s]._graph_out.get("grad_features")
            if g is not None:                                   # the feature-map gradient: fixed size, copied directly
                if self.grad_host[s] is None:
                    self.grad_host[s] = torch.empty((g.numel(),), dtype=g.dtype).pin_memory()
                self.grad_host[s].copy_(g.reshape(-1), non_blocking=True)
                nbytes += g.numel() * g.element_size()
            self.d2h_done[s] = self.s_out.record_event()
        self.d2h_bytes = nbytes

    @property
    def host_views(self) -> List[List[torch.Tensor]]:
        """Per slot: the host copies of the last step's results (views into the pinned staging buffers, flattened), in the
        order of ``RoIPathStep.result_tensors(finalize(...))``. Call after run()."""
        out = []
        for s, slot in enumerate(self.slots):
            cnt, offs, views = self.counts_host[s].tolist(), self.offsets_host[s].tolist(), []
            for j, (t, ci) in enumerate(self.specs[s]):
                rows = int(t.shape[0]) if t.dim() else 1
                live = rows if ci is None else min(max(cnt[ci], 0), rows)
                nbytes = live * (t.numel() // max(rows, 1)) * t.element_size()
                views.append(self.packed_host[s][offs[j]: offs[j] + nbytes].view(t.dtype))
            if self.grad_host[s] is not None:
                views.append(self.grad_host[s])
            out.append(views)
        return out

    def run(self, pinned, steps: int) -> None:
        """`steps` end-to-end steps; returns when every result is on the host. pinned: a dict of host tensors that
        is written into the pinned staging buffers before every H2D copy, or None to send the staging buffers
        as they are (filled beforehand with load_inputs)."""
        for st in (self.s_in, self.s_c, self.s_out, self.s_p):
            st.wait_stream(torch.cuda.current_stream())
        self._h2d(0, pinned)
        for n in range(steps):
            if n + 1 < steps:
                self._h2d(n + 1, pinned)
            self._compute(n)
            if n > 0:
                self._d2h(n - 1)
        self._d2h(steps - 1)
        self.s_out.synchronize()
        for st in (self.s_in, self.s_c, self.s_out, self.s_p):
            torch.cuda.current_stream().wait_stream(st)
