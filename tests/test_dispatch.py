"""coin_b200.patch(): torchvision's CUDA dispatch keys for roi_align / _roi_align_backward / nms are served by libcoinops
(SURVEY.md 8b): an unmodified torchvision / detectron2 caller runs this library's kernels."""
import pytest
import torch
import torchvision

import coin_b200
from coin_b200 import _lib

OPS = ("torchvision::roi_align", "torchvision::_roi_align_backward", "torchvision::nms")


def _cuda_line(op):
    return next(l for l in torch._C._dispatch_dump(op).splitlines() if l.startswith("CUDA:"))


def test_patch_registers_and_unpatch_restores_cuda_kernels():
    before = {op: _cuda_line(op) for op in OPS}
    coin_b200.patch()
    coin_b200.patch()   # idempotent
    try:
        assert coin_b200.is_patched()
        for op in OPS:
            assert "coin_b200/dispatch.py" in _cuda_line(op), _cuda_line(op)
        # the CPU key is untouched: torchvision's CPU kernels still answer CPU tensors
        b = torch.tensor([[0., 0., 10., 10.], [1., 1., 11., 11.], [50., 50., 60., 60.]])
        assert torchvision.ops.nms(b, torch.tensor([0.9, 0.8, 0.7]), 0.5).tolist() == [0, 2]
    finally:
        coin_b200.unpatch()
    assert not coin_b200.is_patched()
    assert {op: _cuda_line(op) for op in OPS} == before


@pytest.fixture
def patched():
    coin_b200.patch()
    yield
    coin_b200.unpatch()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_torchvision_roi_align_runs_this_library(dev, patched, dtype):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 37, 75, generator=g)
    xy = torch.rand(300, 2, generator=g) * torch.tensor([1100.0, 500.0])
    wh = torch.rand(300, 2, generator=g) * 300 + 4
    rois = torch.cat((torch.randint(0, 2, (300, 1), generator=g).float(), xy, xy + wh), dim=1)
    gout = torch.randn(300, 64, 14, 14, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = torchvision.ops.roi_align(xr, rois, (14, 14), 1.0 / 16, 0, True)          # CPU key: torchvision's kernel
    ref.backward(gout)

    xd = x.to(dev, dtype).requires_grad_(True)
    n0 = _lib.lib.coin_launch_count()
    out = torchvision.ops.roi_align(xd, rois.to(dev, dtype), (14, 14), 1.0 / 16, 0, True)
    n1 = _lib.lib.coin_launch_count()
    out.backward(gout.to(dev, dtype))
    n2 = _lib.lib.coin_launch_count()
    assert n1 > n0 and n2 > n1, "torchvision.ops.roi_align did not reach libcoinops"
    assert out.dtype == dtype and xd.grad.dtype == dtype
    if dtype == torch.float32:
        torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5 * float(x.abs().max()))
        torch.testing.assert_close(xd.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-5 * float(xr.grad.abs().max()))
    else:   # fp16 I/O: the reference under autocast; inputs rounded to fp16, arithmetic fp32
        xh = x.half().float().requires_grad_(True)
        refh = torchvision.ops.roi_align(xh, rois.half().float(), (14, 14), 1.0 / 16, 0, True)
        torch.testing.assert_close(out.float().cpu(), refh, rtol=2e-3, atol=2e-3)

    # under autocast torchvision's own wrapper casts to fp32 and re-dispatches to the (patched) CUDA key
    with torch.autocast("cuda", dtype=torch.float16):
        n3 = _lib.lib.coin_launch_count()
        o2 = torchvision.ops.roi_align(xd.detach(), rois.to(dev), (7, 7), 1.0 / 16, 0, True)
        assert _lib.lib.coin_launch_count() > n3
    assert o2.shape == (300, 64, 7, 7)


@pytest.mark.gpu
def test_torchvision_nms_and_batched_nms_run_this_library(dev, patched):
    g = torch.Generator().manual_seed(9)
    xy = torch.rand(5000, 2, generator=g) * 500
    boxes = torch.cat((xy, xy + torch.rand(5000, 2, generator=g) * 120 + 2), dim=1)
    scores = torch.rand(5000, generator=g)
    idxs = torch.randint(0, 8, (5000,), generator=g)
    ref = torchvision.ops.nms(boxes, scores, 0.5)
    refb = torchvision.ops.batched_nms(boxes, scores, idxs, 0.5)
    n0 = _lib.lib.coin_launch_count()
    keep = torchvision.ops.nms(boxes.to(dev), scores.to(dev), 0.5)
    n1 = _lib.lib.coin_launch_count()
    keepb = torchvision.ops.batched_nms(boxes.to(dev), scores.to(dev), idxs.to(dev), 0.5)
    assert n1 > n0 and _lib.lib.coin_launch_count() > n1
    assert torch.equal(keep.cpu(), ref)
    assert torch.equal(keepb.cpu(), refb)
    e = torchvision.ops.nms(torch.empty(0, 4, device=dev), torch.empty(0, device=dev), 0.5)
    assert e.dtype == torch.int64 and e.numel() == 0
