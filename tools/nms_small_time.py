import sys, torch
sys.path.insert(0, '/root/repo')
from coin_b200 import ops, _lib
from coin_b200.sweep import _boxes, _time
dev = torch.device('cuda:0')
for n in (1500, 3000, 5000, 8000):
    b, s, i = _boxes(n, n)
    bd, sd, idd = b.to(dev), s.to(dev), i.to(dev)
    for m in (1024, 1 << 30):
        with _lib.options(COIN_NMS_SEG_MIN=m):
            us = _time(lambda: ops.batched_nms(bd, sd, idd, 0.5, "auto", -1, sync=False), 20)
        print(n, "seg" if m == 1024 else "dense", round(us, 1))
