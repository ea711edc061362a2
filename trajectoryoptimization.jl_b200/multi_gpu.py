"""Multi-GPU plumbing of the hot path.

Problem instances never interact (nothing in the reference couples two ``Problem``s), so the batch is sharded in
contiguous slices, one process per GPU, every buffer local.  The ONLY collective is the all-reduce of the 2-element
vector {sum of merits, max constraint violation} that a solver's global progress report reads -- SUM on the first
entry, MAX on the second -- issued through ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) on
the handle's own device buffer, with no host round trip.
"""
import ctypes as C

import numpy as np

from . import _capi as K
from . import api as TO


def shard_slice(total, rank, world):
    """contiguous slice [lo, hi) of ``total`` instances owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevPtr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def merit_device_tensor(prob, device):
    """zero-copy torch view of the handle's {sum J, max violation} device buffer."""
    import torch
    ptr = C.c_void_p()
    K.check(prob._lib, prob._h, prob._lib.to_merit_device_ptr(prob._h, C.byref(ptr)))
    return torch.as_tensor(_DevPtr(ptr.value, 2), device=device)


def all_reduce_merit(t2, group=None, scratch=None):
    """in-place reduction of a 2-element tensor over the ranks: SUM on [0], MAX on [1] -- ONE collective (an all-gather of the
    2-vectors, reduced locally) instead of a SUM and a MAX all-reduce: the exchange is launch-latency bound (16 bytes per rank).
    ``scratch``: a preallocated [world, 2] tensor on the same device (avoids an allocation per call)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        buf = scratch if scratch is not None else torch.empty((world, 2), dtype=t2.dtype, device=t2.device)
        dist.all_gather_into_tensor(buf, t2.reshape(1, 2), group=group) if t2.is_cuda else dist.all_gather(list(buf.unbind(0)), t2, group=group)
        t2[0] = buf[:, 0].sum()
        t2[1] = buf[:, 1].max()
    return t2


def global_merit(prob, group=None, device_tensor=None):
    """{sum of merits, max violation} over every rank's shard.  With ``device_tensor`` (from merit_device_tensor) the
    reduction stays on the device; otherwise it goes through host arrays (CPU / gloo)."""
    import torch
    if device_tensor is not None:
        # reduce behind whatever is still in flight and hand the result to the current torch stream through an event
        # (to_reduce_merit_async): the handle's own stream keeps running the next iteration
        K.check(prob._lib, prob._h, prob._lib.to_reduce_merit_async(prob._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return all_reduce_merit(device_tensor, group)
    t2 = torch.tensor([float(np.sum(TO.merit(prob))), float(np.max(TO.max_violation(prob)))], dtype=torch.float64)
    return all_reduce_merit(t2, group)
