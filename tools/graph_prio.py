"""Diagnostic: priorities of the kernel nodes of the captured step graph (do graph nodes keep their stream's priority?)."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import pipeline, synth  # noqa: E402
from cuda.bindings import driver as drv  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES["foggy_roi_head"]
step = pipeline.RoIPathStep(shape, dev)
d = step.to_device(synth.image_batch(shape))
step.capture(d, backward=True, keep_graph=True)
graph = drv.CUgraph(step._graph.raw_cuda_graph())
err, _, n = drv.cuGraphGetNodes(graph, 0)
err, nodes, n = drv.cuGraphGetNodes(graph, n)
print("nodes", n, err)
hist = collections.Counter()
for node in nodes:
    err, ty = drv.cuGraphNodeGetType(node)
    if ty != drv.CUgraphNodeType.CU_GRAPH_NODE_TYPE_KERNEL:
        hist[("non-kernel", str(ty))] += 1
        continue
    err, val = drv.cuGraphKernelNodeGetAttribute(node, drv.CUlaunchAttributeID.CU_LAUNCH_ATTRIBUTE_PRIORITY)
    prio = val.priority if err == drv.CUresult.CUDA_SUCCESS else str(err)
    err2, params = drv.cuGraphKernelNodeGetParams(node)
    name = "?"
    if err2 == drv.CUresult.CUDA_SUCCESS:
        try:
            e3, nm = drv.cuFuncGetName(params.func)
            name = (nm.decode() if isinstance(nm, bytes) else str(nm))[:48]
        except Exception as e:  # noqa: BLE001
            name = "func?"
    else:
        name = str(err2)
    hist[(name, prio)] += 1
for k, v in sorted(hist.items(), key=lambda kv: str(kv[0])):
    print(v, k)
