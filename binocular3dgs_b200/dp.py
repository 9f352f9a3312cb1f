"""View-parallel data parallelism for the rasterizer hot path.

The reference has no collective of any kind: its multi-GPU story is one independent
``train.py`` per scene per GPU (``script/run_llff.py:21-37,61-98``).  A rasterizer call
is a pure function of (Gaussians, camera), so the path shards by VIEW: every rank holds
a replica of the Gaussians, renders its own camera, and the per-Gaussian gradients are
summed with ONE in-place all-reduce over a flat FP32 bucket.

``GradientBucket`` is that bucket.  Its segments are laid out
``[means3D 3 | sh 3M | opacity 1 | scales 3 | rotations 4]`` × P as five contiguous
arrays, and the tensors handed out by :meth:`views` alias it, so when the backward of
the operator writes its outputs *into* those views (``Backend.grad_sink``) there is no
pack/copy step between the backward kernel and NCCL: the kernel's stores are the
collective's send buffer.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

SEGMENTS = ("means3D", "shs", "opacities", "scales", "rotations")


class GradientBucket:
    def __init__(self, P: int, M: int, device, dtype=torch.float32):
        self.P, self.M = int(P), int(M)
        self.widths = {"means3D": 3, "shs": 3 * self.M, "opacities": 1, "scales": 3, "rotations": 4}
        self.floats_per_gaussian = sum(self.widths.values())  # 11 + 3M
        # every segment starts on a 16-byte boundary (the backward kernel stores float4s into
        # the rotation segment); the few padding floats stay zero and ride along in the reduce
        offsets, off = {}, 0
        for name in SEGMENTS:
            offsets[name] = off
            off += (self.P * self.widths[name] + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=dtype, device=device)
        self._views: Dict[str, torch.Tensor] = {}
        for name in SEGMENTS:
            n = self.P * self.widths[name]
            v = self.flat[offsets[name]:offsets[name] + n]
            shape = (self.P, self.M, 3) if name == "shs" else (self.P, self.widths[name])
            self._views[name] = v.view(shape)

    def views(self) -> Dict[str, torch.Tensor]:
        """Tensors aliasing the bucket, one per parameter group."""
        return self._views

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def load(self, grads: Dict[str, torch.Tensor]):
        """Copy gradients in (only needed when the backward did not write in place)."""
        for name in SEGMENTS:
            g = grads[name]
            if g.data_ptr() != self._views[name].data_ptr():
                self._views[name].copy_(g.reshape(self._views[name].shape))

    def all_reduce(self, group=None, average: bool = True, async_op: bool = False):
        """Sum (or average) the bucket over the ranks of ``group``; in place."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average and not async_op:
            self.flat.mul_(1.0 / dist.get_world_size(group))
        return work


class PeerGradientBucket(GradientBucket):
    """The same bucket in SYMMETRIC memory, reduced by this library's own kernel.

    ``flat`` is allocated with ``torch.distributed._symmetric_memory`` so that every rank
    holds the device pointers of all peers' buckets over NVLink / NVSwitch.  The backward
    kernel writes into it exactly as into :class:`GradientBucket` (``Backend.grad_sink``);
    :meth:`all_reduce` is then barrier -> ``b3gs_peer_allreduce`` (one in-place two-shot
    kernel: each rank reduces its slice from all peers and stores the sum to all peers) ->
    barrier, all on the current stream.  No NCCL call is on this path.  All ranks of ``group``
    must construct the bucket collectively.  CUDA only.
    """

    def __init__(self, P: int, M: int, device, group=None, dtype=torch.float32):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem
        from . import _backend
        if dtype != torch.float32:
            raise RuntimeError("PeerGradientBucket is float32 only")
        super().__init__(P, M, "meta")              # sizes and offsets only
        group = dist.group.WORLD if group is None else group
        n = (self.flat.numel() + 3) // 4 * 4
        self.flat = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.flat.zero_()
        self._handle = symm_mem.rendezvous(self.flat, group)
        self.world, self.rank = self._handle.world_size, self._handle.rank
        if self.world > 8:
            raise RuntimeError("b3gs_peer_allreduce supports up to 8 peers (one NVSwitch domain)")
        self._ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self._handle.buffer_ptrs])
        self._fn = _backend.native().lib.b3gs_peer_allreduce
        self._fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t,
                             ctypes.c_float, ctypes.c_void_p]
        self._fn.restype = ctypes.c_int
        # NVLS: reduce inside the NVSwitch when the buffer has a multicast mapping
        # (B3GS_DP_MULTIMEM=0 keeps the plain peer loads/stores)
        import os
        self._mc_ptr = 0
        # measured on 8x B200 (18.4 MB bucket): multimem 70 us, plain peer 74 us, NCCL 108 us; on
        # 2x B200: multimem 66 us, plain 46 us, NCCL 64 us -> the switch reduction from 4 ranks up
        mm = os.environ.get("B3GS_DP_MULTIMEM", "auto")
        if mm == "1" or (mm == "auto" and self.world >= 4):
            try:
                self._mc_ptr = int(self._handle.multicast_ptr or 0)     # 0 when the system has no NVLS
            except Exception:
                self._mc_ptr = 0
        self._fn_mc = _backend.native().lib.b3gs_peer_allreduce_multimem
        self._fn_mc.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float,
                                ctypes.c_void_p]
        self._fn_mc.restype = ctypes.c_int
        off = 0
        for name in SEGMENTS:
            cnt = self.P * self.widths[name]
            shape = (self.P, self.M, 3) if name == "shs" else (self.P, self.widths[name])
            self._views[name] = self.flat[off:off + cnt].view(shape)
            off += (cnt + 3) // 4 * 4

    def all_reduce(self, group=None, average: bool = True, async_op: bool = False):
        if async_op:
            raise NotImplementedError("the peer all-reduce is stream-ordered; async_op has no meaning here")
        if self.world == 1:
            return None
        dev = self.flat.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self._handle.barrier(channel=0)        # every peer's backward has written its bucket
            scale = (1.0 / self.world) if average else 1.0
            if self._mc_ptr:
                rc = self._fn_mc(self.world, self.rank, self._mc_ptr, self.flat.numel(), scale, stream)
            else:
                rc = self._fn(self.world, self.rank, self._ptrs, self.flat.numel(), scale, stream)
            if rc != 0:
                raise RuntimeError(f"b3gs_peer_allreduce failed ({rc})")
            self._handle.barrier(channel=1)        # every peer's slice has landed everywhere
        return None


def make_bucket(P: int, M: int, device, group=None, prefer_peer: bool = True):
    """``PeerGradientBucket`` when the process group spans several CUDA ranks and symmetric
    memory can be set up, else the NCCL/gloo ``GradientBucket``.  Returns (bucket, kind)."""
    if (prefer_peer and dist.is_initialized() and dist.get_world_size(group) > 1
            and torch.device(device).type == "cuda"):
        try:
            return PeerGradientBucket(P, M, device, group), "peer"
        except Exception as ex:      # no P2P / symmetric memory on this system: say so, use NCCL
            import warnings
            warnings.warn(f"symmetric-memory bucket unavailable ({ex!r}); falling back to the NCCL all-reduce")
    return GradientBucket(P, M, device), "nccl"


def shard_views(num_views: int, rank: int, world_size: int):
    """Indices of the views rank ``rank`` renders this step: view v goes to rank v % N."""
    return list(range(rank, num_views, world_size))


def reduce_densify_stats(grad_norm: torch.Tensor, visible: torch.Tensor, max_radii: torch.Tensor, group=None):
    """Statistics that must NOT be derived from the reduced gradient:
    the densifier uses the per-view NORM of dL/dmean2D and a visibility count
    (scene/gaussian_model.py:409-411) and the max screen radius (train.py:178).  Each
    rank contributes its own norm / count, summed; radii are max-reduced."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return grad_norm, visible, max_radii
    packed = torch.stack([grad_norm.reshape(-1).float(), visible.reshape(-1).float()], dim=0)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    r = max_radii.clone()
    dist.all_reduce(r, op=dist.ReduceOp.MAX, group=group)
    return packed[0].reshape(grad_norm.shape), packed[1].reshape(visible.shape), r
