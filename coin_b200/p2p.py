"""Gradient all-reduce over NVLink peer memory (BASELINE.json configs[3]): the host side of ``coin_p2p_all_reduce``.

``PeerAllReduce(nelem)`` allocates this rank's fp32 gradient buffer and flag words, shares them with the other ranks of the
node through CUDA IPC (handles travel over ``torch.distributed.all_gather_object``) and maps theirs. ``all_reduce()`` then
sums the buffers in place with five small launches on the given stream - no NCCL kernel, CTAs sized to run beside the
ROIAlign grids (see coin_b200/csrc/p2p_allreduce.cu). One process per GPU, world size <= 8, one NVSwitch domain.
"""
import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from ._lib import check, lib

_FLAG_WORDS = 3 * 8


class PeerAllReduce:
    def __init__(self, nelem: int, device: Optional[torch.device] = None, group=None):
        if not dist.is_initialized():
            raise RuntimeError("coin_b200.p2p: torch.distributed is not initialised")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("coin_b200.p2p: at most 8 ranks (one NVSwitch domain)")
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        quantum = 4 * self.world
        self.nelem = (int(nelem) + quantum - 1) // quantum * quantum
        # own allocations (cudaMalloc'ed by the caching allocator: shareable through CUDA IPC)
        self.buffer = torch.zeros((self.nelem,), dtype=torch.float32, device=self.device)
        self._flags = torch.zeros((_FLAG_WORDS,), dtype=torch.int32, device=self.device)
        self.err = torch.zeros((1,), dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        mine = (self.buffer.untyped_storage()._share_cuda_(), self._flags.untyped_storage()._share_cuda_())
        everyone: List[object] = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._peers = []          # keeps the mapped storages alive
        data_ptrs, flag_ptrs = [], []
        for r, (hd, hf) in enumerate(everyone):
            if r == self.rank:
                data_ptrs.append(self.buffer.data_ptr())
                flag_ptrs.append(self._flags.data_ptr())
                continue
            # open the peer's allocation in THIS rank's device context (the handle's first field is the exporter's device
            # index; opened there, the mapping would not be peer-accessible from our kernels)
            me = self.device.index if self.device.index is not None else torch.cuda.current_device()
            sd = torch.UntypedStorage._new_shared_cuda(me, *hd[1:])
            sf = torch.UntypedStorage._new_shared_cuda(me, *hf[1:])
            self._peers.append((sd, sf))
            data_ptrs.append(sd.data_ptr())
            flag_ptrs.append(sf.data_ptr())
        self._data_arr = (ctypes.c_void_p * self.world)(*data_ptrs)
        self._flag_arr = (ctypes.c_void_p * self.world)(*flag_ptrs)
        self._epoch = 0
        dist.barrier(group=group)   # every rank has mapped every buffer before anyone starts reducing

    def all_reduce(self, offset: int = 0, nelem: Optional[int] = None, max_ctas: int = 0, stream: Optional[torch.cuda.Stream] = None):
        """In-place sum of buffer[offset : offset + nelem] over the ranks (asynchronous on ``stream``, default: current).
        Every rank must make the same sequence of calls."""
        nelem = self.nelem - offset if nelem is None else int(nelem)
        self._epoch += 1
        s = stream if stream is not None else torch.cuda.current_stream()
        check(lib.coin_p2p_all_reduce(self._data_arr, self._flag_arr, self.rank, self.world, int(offset), nelem, self._epoch,
                                      ctypes.c_void_p(self.err.data_ptr()), int(max_ctas), ctypes.c_void_p(s.cuda_stream)))

    def check(self) -> None:
        """Raises if a barrier of an earlier call gave up (a peer never arrived). Synchronises."""
        code = int(self.err.item())
        if code:
            raise RuntimeError(f"coin_b200.p2p: barrier {code - 1} timed out waiting for a peer")
