"""Builds coin_b200/libcoinops.so (sm_100a only) with nvcc. In-tree, no JIT cache.

    python coin_b200/build.py [--force] [--verbose]      (or __graft_entry__.build())

-fmad=false: the integer-valued results of this path (match indices, labels, keep lists) are decided
by float compares; the CPU oracle executes un-fused multiply/add, so the kernels must too.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcoinops.so")
SOURCES = ["capi.cu", "roi_align.cu", "roi_align_sep.cu", "roi_align_reg.cu", "box_codec.cu", "iou_match.cu", "nms.cu", "fusion_nms.cu",
           "det_postprocess.cu", "match_abc.cu", "step_dev.cu", "rpn_proposals.cu", "sampling_loss.cu", "voc_eval.cu", "p2p_allreduce.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-diag-suppress", "128",
    "-Wno-deprecated-gpu-targets",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "roi_common.cuh"), os.path.join(HERE, "..", "include", "coinops.h")]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    # the image exports CC/CXX=/opt/gcc/bin/*, wrappers nvcc cannot drive: use the PATH compiler
    env.pop("CC", None), env.pop("CXX", None)
    objs, procs = [], []
    for src in srcs:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", "g++", "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, env=env)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    if force or procs or _stale(SO, objs):
        cmd = [_nvcc(), "-shared", "-ccbin", "g++", "-cudart", "static", "-o", SO, *objs]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, env=env)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
