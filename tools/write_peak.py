"""What a pure-WRITE stream reaches on this GPU, next to the copy figure of MEASURED_PEAKS.json (read + write bytes of b.copy_(a)).
The ROIAlign forward is a write stream (1.19 GB written, 0.04 GB read from DRAM per launch), the backward a read stream
(1.28 GB read): their fractions of the copy peak are read against these one-directional ceilings in profiles/.
CUDA events, best and median of 20 after 3 warm-ups; the buffers (1.2 GB) exceed the 126 MB L2."""
import json

import torch


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    n = 1536 * 1024 * 196                      # the pooled tensor of the bench shape, fp32
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty(n, dtype=torch.float32, device=dev)
    nbytes = n * 4
    out = {"bytes": nbytes}
    for name, fn, moved in (("fill_kernel_write", lambda: a.fill_(1.0), nbytes),
                            ("memset_write", lambda: a.zero_(), nbytes),
                            ("sum_read", lambda: a.sum(), nbytes),
                            ("copy_read_plus_write", lambda: b.copy_(a), 2 * nbytes)):
        best, med = timed(fn)
        out[name] = {"best_gbs": moved / best / 1e9, "median_gbs": moved / med / 1e9, "best_us": best * 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
