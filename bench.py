#!/usr/bin/env python
"""RoI-path benchmark (BASELINE.json metric: RoI-path images/sec; per-kernel HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the RoI path (coin_b200/pipeline.py: T1-T4, S1-S5) over one batch of
synthetic Foggy-Cityscapes-shaped inputs: BASELINE.json configs[1] (3 images 600x1200, 1000 teacher
RoIs, 100 cloud detections, 12000-box RPN NMS, 2000 proposals, 512 RoIs/image ROIAlign 14x14
forward + backward on a [3,1024,37,75] map). One process per GPU, images sharded across ranks, no
collective on the path (weak scaling). Rank 0 prints ONE JSON line.

--impl reference times the reference's CPU path for the same step (the oracle restatement on
torch/torchvision CPU operators: the reference is pure Python and its dependencies are not
installable here) on the host cores of rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "foggy_roi_head"
METRIC = "roi_path_images_per_sec"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="coin_b200", choices=["coin_b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def host_info():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return model


def cpu_step(batch_one_image, anchors, grad, threads):
    """The reference's CPU path for ONE image of the batch (all stages, ROIAlign fwd + bwd included)."""
    from oracle import pipeline_ref
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    pipeline_ref.run(batch_one_image, backward=True, anchors=anchors, grad=grad)
    return time.perf_counter() - t0


def one_image_view(batch, shape, i=0):
    from coin_b200 import synth
    one = synth.Shape(**{**shape.__dict__, "images": 1})
    return {"shape": one, "features": batch["features"][i:i + 1].contiguous(), "images": [batch["images"][i]]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from coin_b200 import synth
    from oracle import pipeline_ref
    shape = synth.SHAPES[args.workload]
    batch = synth.image_batch(shape)
    threads = os.cpu_count() or 1
    one = one_image_view(batch, shape)
    anchors = pipeline_ref.anchors_for(shape)
    grad = pipeline_ref.head_grad(shape)[: shape.rois]
    for _ in range(min(args.warmup, 1)):
        cpu_step(one, anchors, grad, threads)
    times = [cpu_step(one, anchors, grad, threads) for _ in range(args.steps)]
    total = sum(times)
    value = args.steps * 1 / total
    sample = f"{args.steps} steps x 1 image (of the {shape.images}-image batch), every stage incl. ROIAlign fwd+bwd"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": "reference CPU path = oracle port (torch/torchvision CPU "
                       "operators + restated detectron2/COIN Python); runs on rank 0's host cores only"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "cpu_model": host_info()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from coin_b200 import _lib, pipeline, sharding, synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shape = synth.SHAPES[args.workload]
    # every rank owns different images (COIN_BENCH_SEED_OFFSET: reproduce another rank's data on one GPU)
    batch = synth.image_batch(shape, seed=synth.SEED + rank + int(os.environ.get("COIN_BENCH_SEED_OFFSET", "0")))
    step = pipeline.RoIPathStep(shape, dev)
    pinned = step.host_inputs(batch)
    d = step.h2d(pinned)
    torch.cuda.synchronize()

    # ---- the step as ONE CUDA graph: run_static (device-side lengths, no host round trip) captured once,
    #      with an external CUDA-event pair around each of the two ROIAlign kernels inside the graph
    ev = {"fwd": [], "bwd": []}
    launches0 = _lib.lib.coin_launch_count()
    step.run_static(d, backward=True)              # eager once: counts the launches of one step
    torch.cuda.synchronize()
    launches_per_step = _lib.lib.coin_launch_count() - launches0
    step.kernel_events = ev
    step.capture(d, backward=True)                 # graph inputs = the resident tensors `d`
    step.kernel_events = None

    # ---- device-resident throughput: K graph replays back to back, inputs already in HBM ---------------
    for _ in range(max(args.warmup, 3)):
        step.replay()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        out = step.replay()
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    launches = launches_per_step * args.steps
    # kernel durations inside the graph: replay, then read the external event pairs (one sample per replay)
    fwd_ms, bwd_ms = [], []
    for _ in range(args.steps):
        step.replay()
        torch.cuda.synchronize()
        fwd_ms.append(ev["fwd"][0][0].elapsed_time(ev["fwd"][0][1]))
        bwd_ms.append(ev["bwd"][0][0].elapsed_time(ev["bwd"][0][1]))
    k_fwd_step, k_bwd_step = statistics.mean(fwd_ms), statistics.mean(bwd_ms)
    # the same two launches ALONE on the device (nothing else resident): CUDA events on the launching stream
    # around each launch; the 1.23 GB pooled / gradient tensors exceed the 126 MB L2, so every launch is cold
    iso = {"fwd": [], "bwd": []}
    step.time_roi_kernels(d, iso, iters=max(args.steps, 10), warmup=3)
    torch.cuda.synchronize()
    k_fwd = statistics.mean(a.elapsed_time(b) for a, b in iso["fwd"])
    k_bwd = statistics.mean(a.elapsed_time(b) for a, b in iso["bwd"])

    # ---- end to end: every step copies its inputs from pinned host memory, replays the graph, reads the
    #      lengths back and copies the live results to pinned host memory; consecutive steps are double
    #      buffered (H2D of step n+1 and D2H of step n-1 overlap the graph of step n)
    h2d_bytes = step.input_bytes(pinned)
    pipe = pipeline.PipelinedSteps(step, d, backward=True)
    pipe.load_inputs(pinned)          # the step's inputs sit in the pinned staging buffers (a loader's output)
    pinned = None                     # every step below H2D-copies those pinned buffers
    pipe.run(pinned, 4)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.run(pinned, args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    d2h_bytes = pipe.d2h_bytes
    # un-pipelined latency of ONE end-to-end step (H2D -> graph -> lengths -> D2H, nothing overlapped)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(5):
        pipe.run(pinned, 1)
    l1.record()
    barrier()
    ms_e2e_latency = l0.elapsed_time(l1) / 5
    clocks = sampler.stop()   # sampled from the start of the timed region to the end of the end-to-end region

    ms, ms_e2e, k_fwd, k_bwd, k_fwd_step, k_bwd_step = sharding.max_over_ranks(
        [ms, ms_e2e, k_fwd, k_bwd, k_fwd_step, k_bwd_step], device=dev)

    if rank == 0:
        peak, peak_src = peaks()
        n, c, (h, w) = shape.images, shape.channels, shape.feat_hw
        k = n * shape.rois
        out_bytes = k * c * shape.pooled * shape.pooled * 4
        map_bytes = n * c * h * w * 4
        fwd_bytes = out_bytes + map_bytes + k * 20            # output + map read once + rois
        bwd_bytes = out_bytes + 2 * map_bytes + k * 20        # grad_out read + zero-fill and write of grad map
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
        def kern(ms_alone, ms_in_step, nbytes):
            # ms: the launch alone on the device; ms_in_step: the same launch inside the graph-replayed step,
            # where it shares the SMs with the latency-bound kernels of the other streams
            return {"ms": ms_alone, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms_alone / 1e6,
                    "frac": nbytes / ms_alone / 1e6 / peak, "ms_in_step": ms_in_step,
                    "achieved_gbs_in_step": nbytes / ms_in_step / 1e6, "frac_in_step": nbytes / ms_in_step / 1e6 / peak}

        kernels = {"roi_align_fwd_reg_kernel": kern(k_fwd, k_fwd_step, fwd_bytes),
                   "roi_align_bwd_reg_kernel": kern(k_bwd, k_bwd_step, bwd_bytes)}
        dom = max(kernels, key=lambda name: kernels[name]["ms"])
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                    "peak_source": peak_src, "unit": "GB/s", "frac": kernels[dom]["frac"],
                    "traffic": traffic.get(dom),
                    "share_of_step": kernels[dom]["ms_in_step"] / (ms / args.steps),
                    "timing": "achieved/frac: the launch timed alone with CUDA events on its stream (burst peak); "
                              "*_in_step: the same launch inside the graph, overlapped with the other streams",
                    "kernels": kernels}
        line = {"metric": METRIC, "value": sharding.whole_job_rate(n, world, args.steps, ms), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "images_per_step_per_gpu": n, "rois_per_image": shape.rois,
                           "pooled": shape.pooled, "feature_map": [n, c, h, w], "classes": shape.classes,
                           "teacher_rois": shape.teacher_rois, "rpn_pre_nms": shape.rpn_pre_nms,
                           "l2": "per-step working set (2 x 1.23 GB pooled/grad tensors) >> 126 MB L2",
                           "execution": "one CUDA graph per step (sync-free step, device-side lengths); ROIAlign "
                                        "forward/backward overlap the teacher/matching branch on separate streams",
                           "kernel_timing": "ms_in_step: external CUDA-event pairs captured inside the graph around "
                                            "the two ROIAlign kernels, read after each of `steps` extra replays; "
                                            "ms: the same launches alone, events around each launch",
                           "launches_per_step": launches / args.steps},
                "clocks": clocks,
                "e2e": {"value": sharding.whole_job_rate(n, world, args.steps, ms_e2e), "unit": UNIT,
                        "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": ms_e2e / args.steps, "ms_latency_one_step": ms_e2e_latency,
                        "boundary": "every step: pinned host inputs (feature map, RoIs, deltas, scores, cloud "
                                    "detections, RPN boxes) -> device; graph replay; lengths -> host; detections, "
                                    "A/B/C sets, labels, keep lists and the feature-map gradient -> pinned host. "
                                    "Steps are double buffered (H2D of n+1 and D2H of n-1 overlap the graph of n); "
                                    "ms_latency_one_step is the same step with nothing overlapped"},
                "gpu_launches": int(launches),
                "roofline": roofline}
        # CPU baseline: N=1 only, rank 0, bounded sample = one image of the batch through every stage
        if world == 1:
            from oracle import pipeline_ref
            threads = os.cpu_count() or 1
            one = one_image_view(batch, shape)
            anchors = pipeline_ref.anchors_for(shape)
            grad = pipeline_ref.head_grad(shape)[: shape.rois]
            cpu_step(one, anchors, grad, threads)
            times = [cpu_step(one, anchors, grad, threads) for _ in range(3)]
            line["cpu_baseline"] = {"value": 1.0 / statistics.mean(times), "unit": UNIT, "cores": threads,
                                    "kind": "port", "cpu_model": host_info(),
                                    "sample": "3 timed repeats (+1 warm-up) of 1 image of the batch through every "
                                              "stage, ROIAlign 512 RoIs x 1024 ch x 14x14 fwd+bwd included"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
