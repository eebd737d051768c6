#!/usr/bin/env python
"""RoI-path benchmark (BASELINE.json metric: RoI-path images/sec; per-kernel HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Workloads (BASELINE.json `configs`):
  foggy_roi_head  (default, configs[1]/[2]) one pass of the RoI path (coin_b200/pipeline.py: T1-T4, S1-S5) over 3 images
                  600x1200: 1000 teacher RoIs, 100 cloud detections, 12000-box RPN NMS, 2000 proposals, 512 RoIs/image
                  ROIAlign 14x14 forward + backward on a [3,1024,37,75] map.
  bdd_2000        (configs[3]) the same step at the BDD100K shape (600x1067, 7 classes, 2000 RoIs/image); with N > 1 ranks
                  every step also all-reduces a 0.2 GB fp32 gradient bucket set over NCCL (25 MB buckets, the
                  adaptation-training DDP traffic of trainer.py:66-72), overlapped with the step.
  sweep           (configs[4]) operator sweep: ROIAlign / NMS / IoU+Matcher at 1k..100k boxes, 20-class Clipart shape,
                  each with achieved GB/s or pairs/s, roofline fraction and the CPU path's time beside it.
One process per GPU, images sharded across ranks, no collective on the RoI path (weak scaling). Rank 0 prints ONE JSON
line.

--impl reference times the reference's CPU path for the same step (the oracle restatement on torch/torchvision CPU
operators, pinned to the reference's own outputs by tests/test_reference_goldens_cpu.py: the reference is pure Python
and its dependencies are not installable here) on the host cores of rank 0.
"""
import argparse
import importlib.util
import json
import os
import statistics
import sys
import threading
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="coin_b200", choices=["coin_b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD, choices=["foggy_roi_head", "bdd_2000", "sweep"])
    ap.add_argument("--allreduce-mb", type=float, default=None,
                    help="fp32 gradient bytes all-reduced per step (MB); default 200 for bdd_2000 with N > 1, else 0")
    return ap.parse_args()


def load_synth():
    """coin_b200/synth.py (pure torch) loaded by path: the reference arm must not map libcoinops.so."""
    spec = importlib.util.spec_from_file_location("coin_b200_synth", os.path.join(ROOT, "coin_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["coin_b200_synth"] = mod
    spec.loader.exec_module(mod)
    return mod


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of one GPU, polled through NVML every 5 ms while the timed regions run."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.mx, self.err = index, [], set(), None, None
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(
                    pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:      # noqa: BLE001 - the sampler must never take the benchmark down
            self.err = repr(e)

    def start(self):
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            try:
                self.index = int(visible.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "samples": len(self.sm),
               "reasons": sorted(self.reasons)}
        if self.err:
            out["sampler_error"] = self.err
        return out


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


def workload_config(shape, name):
    n, c, (h, w) = shape.images, shape.channels, shape.feat_hw
    pooled_gb = n * shape.rois * c * shape.pooled * shape.pooled * 4 / 1e9
    # identical in both arms (the driver compares the two lines' `config`); arm-specific notes are top-level keys
    return {"workload": name, "images_per_step_per_gpu": n, "rois_per_image": shape.rois, "pooled": shape.pooled,
            "feature_map": [n, c, h, w], "classes": shape.classes, "teacher_rois": shape.teacher_rois,
            "rpn_pre_nms": shape.rpn_pre_nms,
            "l2": f"no flush needed: the per-step working set (2 x {pooled_gb:.2f} GB pooled / gradient tensors) exceeds the "
                  "126 MB L2 many times over"}


def cpu_step(batch_one_image, anchors, grad, threads):
    """The reference's CPU path for ONE image of the batch (all stages, ROIAlign fwd + bwd included)."""
    from oracle import pipeline_ref
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    pipeline_ref.run(batch_one_image, backward=True, anchors=anchors, grad=grad)
    return time.perf_counter() - t0


def one_image_view(synth, batch, shape, i=0):
    one = synth.Shape(**{**shape.__dict__, "images": 1})
    return {"shape": one, "features": batch["features"][i:i + 1].contiguous(), "images": [batch["images"][i]]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    synth = load_synth()
    from oracle import pipeline_ref
    name = WORKLOAD if args.workload == "sweep" else args.workload
    shape = synth.SHAPES[name]
    batch = synth.image_batch(shape)
    threads = os.cpu_count() or 1
    one = one_image_view(synth, batch, shape)
    anchors = pipeline_ref.anchors_for(shape)
    grad = pipeline_ref.head_grad(shape)[: shape.rois]
    warm = max(min(args.warmup, 1), 1)
    for _ in range(warm):
        cpu_step(one, anchors, grad, threads)
    times, t_begin = [], time.perf_counter()
    for _ in range(args.steps):         # ~4 s per image-step on 16 cores: K steps, bounded to ~150 s of CPU work
        times.append(cpu_step(one, anchors, grad, threads))
        if len(times) >= 3 and time.perf_counter() - t_begin > 150.0:
            break
    steps = len(times)
    total = sum(times)
    value = steps * 1 / total
    sample = (f"{steps} timed passes (+{warm} warm-up) over 1 image of the {shape.images}-image batch, every stage incl. "
              f"ROIAlign fwd+bwd ({shape.rois} RoIs x {shape.channels} ch x {shape.pooled}x{shape.pooled})")
    cfg = workload_config(shape, name)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "steps_timed": steps, "ms_per_step": 1e3 * total / steps / 1,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "reference_arm": ("oracle port of the reference's CPU path (torch/torchvision CPU operators + restated "
                              "detectron2/COIN Python, pinned to the reference's own outputs); rank 0's host cores only"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "cpu_model": host_info()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_link(dev, nbytes=64 << 20, reps=8):
    """Pinned host <-> device copy bandwidth of this rank (GB/s): H2D alone, D2H alone, both at once."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    t_h2d, t_d2h, t_both = timed(h2d), timed(d2h), timed(both)
    return {"h2d_gbs": nbytes / t_h2d / 1e9, "d2h_gbs": nbytes / t_d2h / 1e9,
            "bidir_gbs_each": nbytes / t_both / 1e9, "bytes": nbytes}


def run_e2e(pipeline, step, d, pinned, steps, barrier):
    pipe = pipeline.PipelinedSteps(step, d, backward=True)
    pipe.load_inputs(pinned)          # the step's inputs sit in the pinned staging buffers (a loader's output)
    pipe.run(None, 4)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.run(None, steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(5):
        pipe.run(None, 1)
    l1.record()
    barrier()
    return ms, l0.elapsed_time(l1) / 5, pipe.d2h_bytes


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

    if args.workload == "sweep":
        from coin_b200 import sweep
        if rank == 0:
            print(json.dumps(sweep.run(dev, peaks(), host_info())), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    warmup = max(args.warmup, 3)
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

    # ---- optional gradient all-reduce riding along (configs[3]): DDP-style 25 MB fp32 buckets on a side stream
    ar_mb = args.allreduce_mb
    if ar_mb is None:
        ar_mb = 200.0 if (args.workload == "bdd_2000" and world > 1) else 0.0
    buckets, comm_stream, peer_ar = [], None, None
    if ar_mb > 0 and world > 1:
        n_b = max(int(round(ar_mb / 25.0)), 1)
        buckets = [torch.randn(int(25e6 // 4), device=dev) for _ in range(n_b)]
        # HIGH-priority stream: issued from a normal-priority stream the all-reduce does not get onto the SMs while the ROIAlign
        # grids run and adds its full stand-alone time to the step (tools/ar_overlap.py); from a high-priority one ~70 % of it hides
        comm_stream = torch.cuda.Stream(device=dev, priority=-1)
        # the same buckets in a peer-mapped buffer for the NVLink peer-memory all-reduce (coin_b200/p2p.py), measured beside NCCL
        from coin_b200 import p2p
        peer_ar = p2p.PeerAllReduce(n_b * int(25e6 // 4), dev)
        peer_bucket = peer_ar.nelem // n_b // (4 * world) * (4 * world)
        peer_ar.buffer.normal_()

    def allreduce_buckets(impl="nccl"):
        comm_stream.wait_stream(torch.cuda.current_stream())
        if impl == "p2p":
            for b in range(len(buckets)):
                peer_ar.all_reduce(b * peer_bucket, peer_bucket, stream=comm_stream)
            return
        with torch.cuda.stream(comm_stream):
            for b in buckets:
                dist.all_reduce(b)

    def timed_steps(with_comm, impl="nccl"):
        for _ in range(warmup):
            step.replay()
            if with_comm:
                allreduce_buckets(impl)
        if with_comm:
            torch.cuda.current_stream().wait_stream(comm_stream)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            step.replay()
            if with_comm:
                allreduce_buckets(impl)      # bucket k of step n overlaps the graph of step n + 1
        if with_comm:
            torch.cuda.current_stream().wait_stream(comm_stream)
        t1.record()
        barrier()
        return t0.elapsed_time(t1)

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-resident throughput: K graph replays back to back, inputs already in HBM ---------------
    ms_plain = timed_steps(False)
    ms = timed_steps(True, "nccl") if buckets else ms_plain
    launches = launches_per_step * args.steps
    allreduce = None
    if buckets:
        ms_p2p = timed_steps(True, "p2p")

        def alone(impl):     # the same buckets alone: bus bandwidth of the all-reduce on this box (2 (N-1)/N x bytes / time)
            for _ in range(3):
                allreduce_buckets(impl)
            torch.cuda.current_stream().wait_stream(comm_stream)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(10):
                allreduce_buckets(impl)
            torch.cuda.current_stream().wait_stream(comm_stream)
            a1.record()
            barrier()
            return a0.elapsed_time(a1) / 10
        ar_p2p, ar_nccl = alone("p2p"), alone("nccl")
        peer_ar.check()
        nbytes = sum(b.numel() * 4 for b in buckets)
        bus = lambda t: 2 * (world - 1) / world * nbytes / t / 1e6      # noqa: E731
        allreduce = {"bytes_per_step": nbytes, "buckets": len(buckets), "bucket_mb": 25,
                     "impl": "NCCL all-reduce per bucket on a high-priority side stream (this is what `value` includes)",
                     "ms_alone": ar_nccl, "bus_gbs": bus(ar_nccl), "ms_per_step_with": ms / args.steps,
                     "ms_per_step_without": ms_plain / args.steps,
                     "peer_memory_kernels": {"impl": "coin_p2p_all_reduce (coin_b200/p2p.py): two-shot over CUDA-IPC-mapped buffers",
                                             "ms_alone": ar_p2p, "bus_gbs": bus(ar_p2p), "ms_per_step_with": ms_p2p / args.steps}}

    # per-replay statistics: >= 200 replays, each bracketed by its own event pair
    n_stat = max(args.steps, 200)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_stat)]
    for a, b in evs:
        a.record()
        step.replay()
        b.record()
    torch.cuda.synchronize()
    per = sorted(a.elapsed_time(b) for a, b in evs)
    step_stats = {"replays": n_stat, "min_ms": per[0], "median_ms": per[len(per) // 2], "p95_ms": per[int(0.95 * len(per))],
                  "max_ms": per[-1], "note": "each replay bracketed by its own CUDA-event pair (includes the event gaps)"}

    # kernel durations inside the graph: replay, then read the external event pairs (one sample per replay)
    fwd_ms, bwd_ms = [], []
    for _ in range(min(args.steps, 50)):
        step.replay()
        torch.cuda.synchronize()
        fwd_ms.append(sum(a.elapsed_time(b) for a, b in ev["fwd"]))     # the forward may be split into two launches
        bwd_ms.append(ev["bwd"][0][0].elapsed_time(ev["bwd"][0][1]))
    k_fwd_step, k_bwd_step = statistics.mean(fwd_ms), statistics.mean(bwd_ms)
    # the same two launches ALONE on the device (nothing else resident): CUDA events on the launching stream
    # around each launch; the 1.23 GB pooled / gradient tensors exceed the 126 MB L2, so every launch is cold
    iso = {"fwd": [], "bwd": []}
    step.time_roi_kernels(d, iso, iters=20, warmup=3)
    torch.cuda.synchronize()
    k_fwd = statistics.mean(a.elapsed_time(b) for a, b in iso["fwd"])
    k_bwd = statistics.mean(a.elapsed_time(b) for a, b in iso["bwd"])

    # ---- end to end: every step copies its inputs from pinned host memory, replays the graph, reads the
    #      lengths back and copies the live results to pinned host memory; consecutive steps are double
    #      buffered (H2D of step n+1 and D2H of step n-1 overlap the graph of step n)
    h2d_bytes = step.input_bytes(pinned)
    e2e_steps = min(args.steps, 100)
    ms_e2e, ms_e2e_latency, d2h_bytes = run_e2e(pipeline, step, d, pinned, e2e_steps, barrier)
    # the same boundary in the reference's autocast dtype: fp16 feature map in, fp16 gradient out (fp32 arithmetic)
    step16 = pipeline.RoIPathStep(shape, dev, share=step, io_dtype=torch.float16)
    pinned16 = step16.host_inputs(batch)
    d16 = step16.h2d(pinned16)
    torch.cuda.synchronize()
    h2d16 = step16.input_bytes(pinned16)
    ms_e2e16, ms_lat16, d2h16 = run_e2e(pipeline, step16, d16, pinned16, e2e_steps, barrier)
    link = measure_link(dev)
    clocks = sampler.stop()   # sampled from the start of the timed region to the end of the end-to-end regions

    (ms, ms_plain, ms_e2e, ms_e2e16, k_fwd, k_bwd, k_fwd_step, k_bwd_step) = sharding.max_over_ranks(
        [ms, ms_plain, ms_e2e, ms_e2e16, k_fwd, k_bwd, k_fwd_step, k_bwd_step], device=dev)
    link_min = sharding.max_over_ranks([-link["h2d_gbs"], -link["d2h_gbs"], -link["bidir_gbs_each"]], device=dev)

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
        dom = max(kernels, key=lambda name: kernels[name]["ms_in_step"])       # the slowest kernel of the measured step
        step_ms = ms / args.steps
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                    "peak_source": peak_src, "unit": "GB/s", "frac": kernels[dom]["frac"],
                    "frac_in_step": kernels[dom]["frac_in_step"], "achieved_in_step": kernels[dom]["achieved_gbs_in_step"],
                    "traffic": traffic.get(dom),
                    "share_of_step": kernels[dom]["ms_in_step"] / step_ms,
                    "step_bytes_over_hbm_bound": (fwd_bytes + bwd_bytes) / 1e6 / peak / step_ms,
                    "timing": "kernel = the one that takes longest INSIDE the step; achieved/frac: that launch timed alone "
                              "with CUDA events on its stream (burst peak); *_in_step: the same launch inside the graph, "
                              "overlapped with the other streams; step_bytes_over_hbm_bound: (fwd + bwd algorithmic bytes) "
                              "/ peak / ms_per_step",
                    "kernels": kernels}
        cfg = workload_config(shape, args.workload)
        execution = {"mode": "one CUDA graph per step (sync-free step, device-side lengths); ROIAlign forward/backward overlap "
                             "the teacher/matching branch on separate streams",
                     "launches_per_step": launches / args.steps}
        e2e_cfg = {"value": sharding.whole_job_rate(n, world, e2e_steps, ms_e2e), "unit": UNIT,
                   "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                   "ms_per_step": ms_e2e / e2e_steps, "ms_latency_one_step": ms_e2e_latency,
                   "link_bound_ms_per_step": max(h2d_bytes, d2h_bytes) / (-link_min[2]) / 1e6,
                   "boundary": "every step: pinned host inputs (fp32 feature map, RoIs, deltas, scores, cloud "
                               "detections, RPN boxes) -> device; graph replay; lengths -> host; detections, "
                               "A/B/C sets, labels, keep lists and the fp32 feature-map gradient -> pinned host. "
                               "The ~100 variable-length results are packed on the device (coin_pack_rows) and leave in one copy. "
                               "Steps are double buffered (H2D of n+1 and D2H of n-1 overlap the graph of n); "
                               "ms_latency_one_step is the same step with nothing overlapped; link_bound_ms_per_step "
                               "= max(h2d, d2h bytes) / measured bidirectional link GB/s",
                   "fp16_boundary": {"value": sharding.whole_job_rate(n, world, e2e_steps, ms_e2e16), "unit": UNIT,
                                     "h2d_bytes_per_step": h2d16, "d2h_bytes_per_step": d2h16,
                                     "ms_per_step": ms_e2e16 / e2e_steps, "ms_latency_one_step": ms_lat16,
                                     "link_bound_ms_per_step": max(h2d16, d2h16) / (-link_min[2]) / 1e6,
                                     "note": "feature map and its gradient cross the boundary in fp16, the reference's "
                                             "autocast dtype (trainer.py:175,187); arithmetic and the pooled tensor stay fp32"},
                   "link": {"h2d_gbs": -link_min[0], "d2h_gbs": -link_min[1], "bidir_gbs_each": -link_min[2],
                            "note": "pinned 64 MB copies, slowest rank; bidir = both directions at once, per direction"}}
        line = {"metric": METRIC, "value": sharding.whole_job_rate(n, world, args.steps, ms), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": warmup, "ms_per_step": step_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "execution": execution, "clocks": clocks, "step_stats": step_stats, "e2e": e2e_cfg,
                "gpu_launches": int(launches), "roofline": roofline}
        if allreduce:
            line["allreduce"] = allreduce
        # CPU baseline: N=1 only, rank 0, bounded sample = one image of the batch through every stage
        if world == 1:
            from oracle import pipeline_ref
            threads = os.cpu_count() or 1
            one = one_image_view(synth, batch, shape)
            anchors = pipeline_ref.anchors_for(shape)
            grad = pipeline_ref.head_grad(shape)[: shape.rois]
            cpu_step(one, anchors, grad, threads)
            times = [cpu_step(one, anchors, grad, threads) for _ in range(3)]
            line["cpu_baseline"] = {"value": 1.0 / statistics.mean(times), "unit": UNIT, "cores": threads,
                                    "kind": "port", "cpu_model": host_info(),
                                    "sample": f"3 timed repeats (+1 warm-up) of 1 image of the batch through every stage, ROIAlign "
                                              f"{shape.rois} RoIs x {c} ch x {shape.pooled}x{shape.pooled} fwd+bwd included"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
