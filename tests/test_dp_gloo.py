"""Host-side logic of the view-parallel step on CPU: world_size-2 gloo processes.
The bucket all-reduce must equal the sum (or mean) of the per-rank gradients, the
tensors handed out must alias the flat buffer (no pack copy), view sharding must be a
partition, and densify statistics must be reduced as norms / maxima, not via the
reduced gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from binocular3dgs_b200.dp import SEGMENTS, GradientBucket, reduce_densify_stats, shard_views
    P, M = 257, 4
    bucket = GradientBucket(P, M, device="cpu")
    # P = 257 is odd: every segment still starts on a 16-byte boundary (<= 3 padding floats each)
    assert bucket.floats_per_gaussian == 11 + 3 * M and P * 23 * 4 <= bucket.nbytes <= (P * 23 + 15) * 4
    assert all((v.data_ptr() - bucket.flat.data_ptr()) % 16 == 0 for v in bucket.views().values())
    views = bucket.views()
    def make_local(r):
        g = torch.Generator().manual_seed(100 + r)
        return {n: torch.randn(views[n].shape, generator=g) for n in SEGMENTS}

    local = make_local(rank)
    total = {n: sum(make_local(r)[n] for r in range(world)) for n in SEGMENTS}
    # (a) in-place path: write straight into the views (what Backend.grad_sink does)
    for n in SEGMENTS:
        views[n].copy_(local[n])
        assert views[n].data_ptr() >= bucket.flat.data_ptr()
    bucket.all_reduce(average=False)
    ok = True
    for n in SEGMENTS:
        ok &= torch.allclose(views[n], total[n], atol=1e-6)
    # (b) copy-in path + averaging
    bucket.load(local)
    bucket.all_reduce(average=True)
    for n in SEGMENTS:
        ok &= torch.allclose(views[n], total[n] / world, atol=1e-6)
    # (c) view sharding is a partition
    mine = shard_views(7, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok &= sorted(sum(gathered, [])) == list(range(7))
    # (d) densify statistics: norms are summed, radii maxed
    norm = torch.full((P, 1), float(rank + 1))
    visible = torch.ones(P, 1) * (rank == 0)
    radii = torch.full((P,), rank * 10, dtype=torch.int32)
    n2, v2, r2 = reduce_densify_stats(norm, visible, radii)
    ok &= bool((n2 == sum(range(1, world + 1))).all()) and bool((v2 == 1).all()) and bool((r2 == (world - 1) * 10).all())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_bucket_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_bucket_single_process_is_noop():
    from binocular3dgs_b200.dp import GradientBucket
    b = GradientBucket(10, 16, "cpu")
    assert b.floats_per_gaussian == 59
    b.views()["scales"].fill_(2.0)
    assert b.all_reduce() is None
    assert float(b.flat.sum()) == 2.0 * 30


@pytest.mark.gpu
def test_peer_all_reduce_matches_nccl_on_two_gpus():
    """dp.PeerGradientBucket (symmetric memory + b3gs_peer_allreduce) against the NCCL
    all-reduce, and bit-identity of the replicas — tools/peer_check.py under torchrun.
    Needs two GPUs (skipped on a single-GPU box; run at N = 2 and N = 8 on B200 this round)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          os.path.join(root, "tools", "peer_check.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("replicas bit-identical") == 3


def test_make_bucket_without_cuda_is_the_plain_bucket():
    """make_bucket only chooses the symmetric-memory bucket for CUDA devices inside an
    initialised multi-rank group; everywhere else it is the plain (gloo / NCCL) bucket."""
    from binocular3dgs_b200.dp import GradientBucket, make_bucket
    b, kind = make_bucket(33, 4, "cpu")
    assert kind == "nccl" and type(b) is GradientBucket
    assert b.flat.numel() >= 33 * 23 and all((v.data_ptr() - b.flat.data_ptr()) % 16 == 0 for v in b.views().values())
