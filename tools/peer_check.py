"""2..8-GPU check of dp.PeerGradientBucket against the NCCL all-reduce (run under torchrun):
values, bit-identity across ranks, and timing of both exchange paths on the bench's bucket.
usage: torchrun --nproc-per-node N tools/peer_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from binocular3dgs_b200.dp import SEGMENTS, GradientBucket, PeerGradientBucket  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)

for P, M in ((1001, 4), (200_000, 4), (300_001, 16), (1_000_000, 4)):
    peer, ref = PeerGradientBucket(P, M, dev), GradientBucket(P, M, dev)
    g = torch.Generator().manual_seed(7 * rank + P)
    for name in SEGMENTS:
        v = torch.randn(peer.views()[name].shape, generator=g).to(dev)
        peer.views()[name].copy_(v)
        ref.views()[name].copy_(v)
    peer.all_reduce(average=True)
    ref.all_reduce(average=True)
    torch.cuda.synchronize()
    worst = 0.0
    for name in SEGMENTS:
        a, b = peer.views()[name], ref.views()[name]
        worst = max(worst, float((a - b).abs().max() / b.abs().max()))
    # bit-identical replicas: compare a checksum of the raw bits across ranks
    bits = peer.flat.view(torch.int32).to(torch.int64).sum().reshape(1)
    lo, hi = bits.clone(), bits.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert worst < 1e-6, worst
    assert int(lo) == int(hi), "replicas differ"

    def time_it(fn, iters=50):
        for _ in range(10):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    t_peer = time_it(lambda: peer.all_reduce(average=True))
    t_nccl = time_it(lambda: ref.all_reduce(average=True))
    mc = bool(peer._mc_ptr)
    t_plain = t_peer
    if mc:      # also time the plain peer load/store kernel
        keep, peer._mc_ptr = peer._mc_ptr, 0
        t_plain = time_it(lambda: peer.all_reduce(average=True))
        peer._mc_ptr = keep
    # round 1's protocol: the two barriers as host-issued signal-pad kernels around the reduction
    peer._host_barriers = True
    t_host = time_it(lambda: peer.all_reduce(average=True))
    peer._host_barriers = False
    sweep = ""
    if os.environ.get("B3GS_AR_SWEEP") and P >= 200_000:          # tuning: requests in flight x resident blocks
        for u in (2, 4, 8):
            for b in (2, 4, 8):
                os.environ["B3GS_AR_UNROLL"], os.environ["B3GS_AR_BPSM"] = str(u), str(b)
                sweep += " u%d/b%d %.1f" % (u, b, time_it(lambda: peer.all_reduce(average=True), 30) * 1e3)
        del os.environ["B3GS_AR_UNROLL"], os.environ["B3GS_AR_BPSM"]
        if rank == 0:
            print("  sweep (us):" + sweep, flush=True)
    if rank == 0:
        print("P=%d M=%d (%.1f MB) world=%d: rel diff vs NCCL %.2e, replicas bit-identical; fused kernel (%s) %.1f us, "
              "fused plain peer %.1f us, host-issued barriers %.1f us, NCCL %.1f us"
              % (P, M, peer.nbytes / 1e6, world, worst, "multimem" if mc else "peer", t_peer * 1e3, t_plain * 1e3,
                 t_host * 1e3, t_nccl * 1e3), flush=True)
dist.destroy_process_group()
