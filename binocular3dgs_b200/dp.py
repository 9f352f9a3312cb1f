"""View-parallel data parallelism for the rasterizer hot path.

The reference has no collective of any kind: its multi-GPU story is one independent
``train.py`` per scene per GPU (``script/run_llff.py:21-37,61-98``).  A rasterizer call
is a pure function of (Gaussians, camera), so the path shards by VIEW: every rank holds
a replica of the Gaussians, renders its own camera, and the per-Gaussian gradients are
summed with ONE in-place all-reduce over a flat FP32 bucket.

``GradientBucket`` is that bucket.  Its segments are laid out
``[means3D 3 | sh 3M | opacity 1 | scales 3 | rotations 4]`` × P as five contiguous
arrays, and the tensors handed out by :meth:`views` alias it, so when the backward of
the operator writes its outputs *into* those views (``_C.grad_sink = bucket``) there is no
pack/copy step between the backward kernel and the collective: the kernel's stores are
the collective's send buffer.

Which bucket for which trainer
* The tensors fed to the rasterizer are LEAVES (activated parameters held directly, or
  ``_C`` driven without autograd): ``_C.grad_sink = bucket``.  The first backward of a step
  writes into the bucket and autograd adopts those views as ``.grad`` without a copy; a
  second backward in the same step (the binocular pair, train.py:100,128) is added by
  autograd in place — or, without autograd, accumulated by the kernel itself
  (``B3GS_BWD_ACCUMULATE``).  ``bucket.all_reduce()`` ends the step.
* The rasterizer inputs are COMPUTED from raw parameters (the reference's GaussianModel:
  ``exp`` / ``sigmoid`` / ``normalize`` / ``cat``, scene/gaussian_model.py:95-115): what must be
  summed over ranks is the gradient of the RAW parameters, which autograd forms after the
  rasterizer's backward.  Use :class:`ParameterBucket`: it points every parameter's
  ``.grad`` at a slice of one flat buffer before ``loss.backward()``, autograd accumulates
  there in place, and one all-reduce of the flat buffer follows.  A grad sink must NOT be
  used for this case: it would reduce gradients the optimizer never reads.
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

    # ---- gradient-sink protocol (``_C.grad_sink = bucket``; _backend.Backend.rasterize_gaussians_backward)
    _fresh = True
    _backwards = 0
    backwards_per_step = 1      # 2 for the binocular pair (train.py:100,128): which backward is the last

    def begin_step(self):
        """The next backward overwrites the bucket (called by :meth:`all_reduce`)."""
        self._fresh = True
        self._backwards = 0

    def acquire(self, autograd: bool):
        """-> (views or None, accumulate).  First backward of a step: the views, overwritten.
        Later ones: None under autograd (fresh tensors; autograd adds them into the adopted
        views), else the views again with kernel-side accumulation."""
        self._backwards += 1
        if self._fresh:
            self._fresh = False
            return self._views, False
        return (None, False) if autograd else (self._views, True)

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
        """Sum (or average) the bucket over the ranks of ``group``; in place.  Ends the step
        for the gradient-sink protocol."""
        self.begin_step()
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        if average and async_op:
            raise ValueError("average=True needs the result: use async_op=False, or average=False and scale "
                             "after work.wait()")
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average:
            self.flat.mul_(1.0 / dist.get_world_size(group))
        return work


class PeerGradientBucket(GradientBucket):
    """The same bucket in SYMMETRIC memory, reduced by this library's own kernel.

    ``flat`` is allocated with ``torch.distributed._symmetric_memory`` so that every rank
    holds the device pointers of all peers' buckets over NVLink / NVSwitch.  The backward
    kernel writes into it exactly as into :class:`GradientBucket` (``Backend.grad_sink``);
    :meth:`all_reduce` is then ONE launch of ``b3gs_peer_allreduce_fused``: an in-place two-shot
    all-reduce (each rank reduces its slice from all peers and stores the sum to all peers, through
    the NVSwitch from 4 ranks up) with both cross-rank barriers inside the kernel (flag words behind
    the data) and programmatic dependent launch behind the backward's last kernel, on the current
    stream.  No NCCL call is on this path.  All ranks of ``group``
    must construct the bucket collectively.  CUDA only.
    """

    def __init__(self, P: int, M: int, device, group=None, dtype=torch.float32):
        if dtype != torch.float32:
            raise RuntimeError("PeerGradientBucket is float32 only")
        super().__init__(P, M, "meta")              # sizes and offsets only
        self._setup_symmetric((self.flat.numel() + 3) // 4 * 4, device, group)
        off = 0
        for name in SEGMENTS:
            cnt = self.P * self.widths[name]
            shape = (self.P, self.M, 3) if name == "shs" else (self.P, self.widths[name])
            self._views[name] = self.flat[off:off + cnt].view(shape)
            off += (cnt + 3) // 4 * 4

    def _setup_symmetric(self, n: int, device, group):
        import ctypes
        import os

        import torch.distributed._symmetric_memory as symm_mem
        from . import _backend
        group = dist.group.WORLD if group is None else group
        self._group = group
        # 64 flag words behind the data: the in-kernel barriers of b3gs_peer_allreduce_fused
        self._storage = symm_mem.empty(n + 64, dtype=torch.float32, device=device)
        self._storage.zero_()
        self.flat = self._storage[:n]
        self._epoch = 0
        self._handle = symm_mem.rendezvous(self._storage, group)
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
        self._fn_fused = _backend.native().lib.b3gs_peer_allreduce_fused
        self._fn_fused.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p,
                                   ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint, ctypes.c_float, ctypes.c_void_p]
        self._fn_fused.restype = ctypes.c_int
        # B3GS_DP_BARRIERS=host: the two barriers as symmetric-memory signal-pad kernels around the
        # reduction kernel (round 1's path, kept for A/B); default: inside the one fused kernel
        self._host_barriers = os.environ.get("B3GS_DP_BARRIERS", "kernel") == "host"
        # the exchange plan of b3gs_backward_exchange: the last backward of a step runs K8+K9 chunk by
        # chunk and all-reduces every finished chunk while the next is computed.  OFF by default
        # (B3GS_DP_OVERLAP=1 or bucket.overlap = True turns it on): measured on 2x B200 it LOSES —
        # 1M Gaussians, 92 MB bucket: 1.794 ms (2 chunks) / 1.817 (4) / 1.881 (8) against 1.765 ms for
        # backward then one fused exchange; fern pair 1.276 / 1.302 / 1.402 against 1.253 ms.  The
        # exchange is NVLink-bound (613 GB/s per direction = 80 % of the measured 770 GB/s peer peak)
        # and 1.7x longer than the only kernel it may legally overlap (K8+K9, 89 us); what chunking
        # hides is less than what its extra cross-GPU barriers and the HBM contention cost.
        lib = _backend.native().lib
        self._plan = ctypes.c_void_p(0)
        lib.b3gs_exchange_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        lib.b3gs_exchange_create.restype = ctypes.c_int
        lib.b3gs_exchange_epoch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint]
        lib.b3gs_exchange_epoch.restype = ctypes.c_uint
        with torch.cuda.device(device):
            rc = lib.b3gs_exchange_create(self.world, self.rank, self._ptrs, self._mc_ptr or None, n, n,
                                          ctypes.byref(self._plan))
        if rc != 0:
            raise RuntimeError(f"b3gs_exchange_create failed ({rc})")
        self._plan_epoch = lib.b3gs_exchange_epoch
        self.overlap = os.environ.get("B3GS_DP_OVERLAP", "0") == "1" and not self._host_barriers
        self._exchanged = False
        self.average = True
        torch.cuda.synchronize(device)          # flags zeroed everywhere before anyone's first epoch
        dist.barrier(group)

    def exchange_if_last(self):
        """(plan handle, scale) when the backward that just acquired the sink is the last of the step
        and the overlap is on: that backward is then b3gs_backward_exchange and :meth:`all_reduce`
        has nothing left to do.  ``self.average`` decides the scale."""
        if not self.overlap or self.world == 1 or self._backwards != self.backwards_per_step:
            return None
        self._exchanged = True
        return self._plan, (1.0 / self.world) if self.average else 1.0

    def all_reduce(self, group=None, average: bool = True, async_op: bool = False):
        if async_op:
            raise NotImplementedError("the peer all-reduce is stream-ordered; async_op has no meaning here")
        if group is not None and group is not self._group:
            raise ValueError("PeerGradientBucket reduces over the group it was constructed with")
        self.begin_step()
        if self.world == 1:
            return None
        if self._exchanged:         # done inside the last backward (b3gs_backward_exchange)
            self._exchanged = False
            if average != self.average:
                raise ValueError("the exchange fused into the backward used average=%r" % self.average)
            return None
        dev = self.flat.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            scale = (1.0 / self.world) if average else 1.0
            if not self._host_barriers:
                epoch = self._plan_epoch(self._plan, 0, 0) + 1          # one epoch sequence for both entry points
                self._plan_epoch(self._plan, 1, epoch)
                rc = self._fn_fused(self.world, self.rank, self._ptrs, self._mc_ptr or None, self.flat.numel(),
                                    self.flat.numel(), epoch, scale, stream)
                if rc != 0:
                    raise RuntimeError(f"b3gs_peer_allreduce_fused failed ({rc})")
                return None
            self._handle.barrier(channel=0)        # every peer's backward has written its bucket
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


class ParameterBucket:
    """One flat buffer holding ``.grad`` of every parameter of a model, reduced in one call.

    For trainers whose rasterizer inputs are computed from raw parameters (the reference's
    ``GaussianModel``).  Per step::

        bucket.attach()            # zero the buffer, point every p.grad at its slice
        loss.backward()            # autograd accumulates into the slices in place
        bucket.all_reduce()        # one collective over (11 + 3M) floats per Gaussian
        optimizer.step(); optimizer.zero_grad(set_to_none=True)     # as train.py:192-193

    ``attach`` must be called again after every ``zero_grad(set_to_none=True)`` and after the
    densifier replaced the parameters (``rebuild``: scene/gaussian_model.py:334-345 creates new
    ``nn.Parameter`` objects of a new length).  With ``peer=True`` (CUDA, several ranks) the
    buffer lives in symmetric memory and is reduced by ``b3gs_peer_allreduce``; otherwise by
    ``torch.distributed.all_reduce`` (NCCL or gloo).
    """

    def __init__(self, params, group=None, peer: bool = False):
        self.group, self.peer = group, peer
        self.rebuild(params)

    def rebuild(self, params):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("ParameterBucket needs at least one parameter")
        dev, dtype = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dtype for p in self.params):
            raise ValueError("all parameters of a ParameterBucket must share device and dtype")
        self.offsets, off = [], 0
        for p in self.params:                     # 16-byte aligned slices
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self._inner = None
        if self.peer and dev.type == "cuda" and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            self._inner = _PeerFlat(off, dev, self.group)
            self.flat = self._inner.flat
        else:
            self.flat = torch.zeros(off, dtype=dtype, device=dev)
        self.slices = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, self.params)]

    def attach(self):
        self.flat.zero_()
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view(p.shape)    # a fresh view: autograd adds into it in place

    def all_reduce(self, average: bool = True):
        for p, sl in zip(self.params, self.slices):
            if p.grad is None or p.grad.data_ptr() != sl.data_ptr():
                raise RuntimeError("ParameterBucket.all_reduce: a parameter's .grad no longer aliases the bucket "
                                   "(call attach() before backward, after every zero_grad(set_to_none=True))")
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        if self._inner is not None:
            self._inner.all_reduce(average=average)
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if average:
            self.flat.mul_(1.0 / dist.get_world_size(self.group))


class _PeerFlat(PeerGradientBucket):
    """A bare symmetric-memory float buffer reduced by b3gs_peer_allreduce (no segments)."""

    def __init__(self, n_floats: int, device, group=None):
        GradientBucket.__init__(self, 0, 0, "meta")
        self._setup_symmetric((int(n_floats) + 3) // 4 * 4, device, group)


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


def seed_lockstep(seed: int, rank: Optional[int] = None):
    """Generator discipline for view-parallel training through densification.
    ``densify_and_split`` draws the new positions with ``torch.normal`` on CUDA tensors
    (scene/gaussian_model.py:363-364), i.e. from the CUDA default generator: it must produce
    the SAME stream on every rank, or the replicas diverge at the first split.  View and
    binocular-shift sampling use Python ``random`` and the CPU torch generator
    (train.py:92,125-126): they must DIFFER per rank, or all ranks render the same view."""
    import random
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)                 # identical on every rank
    random.seed(seed + 1000003 * (rank + 1))             # rank-distinct
    torch.default_generator.manual_seed(seed + 1000003 * (rank + 1))


def replica_checksum(tensors) -> torch.Tensor:
    """Order-independent-of-nothing, bit-exact digest of a list of tensors: int64 sums of their
    raw 32-bit words and of word * (index + 1).  Equal tensors <=> (overwhelmingly) equal digests."""
    acc = []
    for t in tensors:
        w = t.detach().contiguous().view(-1).view(torch.int32).to(torch.int64)
        idx = torch.arange(1, w.numel() + 1, device=w.device, dtype=torch.int64)
        acc += [w.sum(), (w * (idx % 65521)).sum(), torch.tensor(w.numel(), device=w.device, dtype=torch.int64)]
    return torch.stack(acc)


def replicas_identical(tensors, group=None) -> bool:
    """True iff every rank of ``group`` holds bit-identical ``tensors`` (one small all-gather)."""
    d = replica_checksum(tensors)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return True
    out = [torch.empty_like(d) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, d, group=group)
    return all(bool(torch.equal(o, out[0])) for o in out)
