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
    assert out.stdout.count("replicas bit-identical") == 4


def test_make_bucket_without_cuda_is_the_plain_bucket():
    """make_bucket only chooses the symmetric-memory bucket for CUDA devices inside an
    initialised multi-rank group; everywhere else it is the plain (gloo / NCCL) bucket."""
    from binocular3dgs_b200.dp import GradientBucket, make_bucket
    b, kind = make_bucket(33, 4, "cpu")
    assert kind == "nccl" and type(b) is GradientBucket
    assert b.flat.numel() >= 33 * 23 and all((v.data_ptr() - b.flat.data_ptr()) % 16 == 0 for v in b.views().values())


# ------------------------------------------------------------------ raw-parameter models
def _activations(raw):
    """The reference's parameterisation (scene/gaussian_model.py:27-40, :95-115)."""
    return dict(means3D=raw["xyz"], shs=torch.cat((raw["f_dc"], raw["f_rest"]), dim=1),
                opacities=torch.sigmoid(raw["opacity"]), scales=torch.exp(raw["scaling"]),
                rotations=torch.nn.functional.normalize(raw["rotation"]))


def _toy_render_loss(act, view_seed):
    """A stand-in for render + loss: a fixed, view-dependent, non-linear function of the
    ACTIVATED inputs (the rasterizer itself needs a GPU; what is under test is which gradients
    get reduced)."""
    g = torch.Generator().manual_seed(view_seed)
    loss = 0.0
    for k in sorted(act):
        w = torch.randn(act[k].shape, generator=g)
        loss = loss + (act[k] * w).sum() + 0.1 * ((act[k] * w) ** 2).sum()
    return loss


def _raw_params(P, M, seed=3):
    g = torch.Generator().manual_seed(seed)
    shapes = dict(xyz=(P, 3), f_dc=(P, 1, 3), f_rest=(P, M - 1, 3), opacity=(P, 1), scaling=(P, 3), rotation=(P, 4))
    return {k: torch.nn.Parameter(torch.randn(s, generator=g) * 0.5) for k, s in shapes.items()}


def _densify_step(raw, grad_norm, denom, lockstep_gen, grad_threshold=0.02, min_opacity=-1.0):
    """Restatement of densify_and_prune's data flow (scene/gaussian_model.py:347-407): clone the
    Gaussians whose averaged view-space gradient norm exceeds the threshold, displace the clones
    with torch.normal from the LOCKSTEP generator (:363-364), prune by opacity (:396-404)."""
    avg = grad_norm / denom.clamp_min(1)
    sel = (avg.squeeze(1) >= grad_threshold)
    out = {}
    for k, p in raw.items():
        new = p.data[sel].clone()
        if k == "xyz":
            std = torch.exp(raw["scaling"].data[sel])
            new = new + torch.normal(torch.zeros_like(std), std, generator=lockstep_gen)
        out[k] = torch.cat((p.data, new), dim=0)
    keep = out["opacity"].squeeze(1) > min_opacity
    return {k: v[keep] for k, v in out.items()}, int(sel.sum())


def _raw_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from binocular3dgs_b200 import dp
    P, M = 101, 4
    raw = _raw_params(P, M)
    bucket = dp.ParameterBucket(list(raw.values()))
    ok = True
    # ---- one step: attach -> backward (rank-distinct view) -> all_reduce
    bucket.attach()
    _toy_render_loss(_activations(raw), view_seed=50 + rank).backward()
    for p in raw.values():
        ok &= p.grad.data_ptr() >= bucket.flat.data_ptr()            # autograd accumulated in place
    bucket.all_reduce(average=True)
    # expected: mean over ranks of the RAW-parameter gradients, each computed alone
    for name in raw:
        exp = 0
        for r in range(world):
            rr = _raw_params(P, M)
            _toy_render_loss(_activations(rr), view_seed=50 + r).backward()
            exp = exp + rr[name].grad / world
        ok &= torch.allclose(raw[name].grad, exp, atol=1e-6, rtol=1e-5)
    ok &= dp.replicas_identical([p.grad for p in raw.values()])
    # a detached .grad is an error, not a silent no-op
    list(raw.values())[0].grad = None
    try:
        bucket.all_reduce()
        ok = False
    except RuntimeError:
        pass
    # ---- densification in lockstep: rank-distinct statistics, reduced; identical decisions
    stats_g = torch.Generator().manual_seed(900 + rank)
    norm = torch.rand(P, 1, generator=stats_g) * 0.05          # this rank's ||dL/dmean2D||
    visible = (torch.rand(P, 1, generator=stats_g) > 0.3).float()
    radii = torch.randint(0, 40, (P,), generator=stats_g, dtype=torch.int32)
    n2, v2, r2 = dp.reduce_densify_stats(norm * visible, visible, radii)
    lock = torch.Generator().manual_seed(1234)                  # what dp.seed_lockstep does for the CUDA generator
    new_raw, n_cloned = _densify_step(raw, n2, v2, lock)
    ok &= n_cloned > 0 and new_raw["xyz"].shape[0] == P + n_cloned
    ok &= dp.replicas_identical(list(new_raw.values()))
    # and the converse: a rank-distinct generator breaks the replicas (the check can fail)
    bad_raw, _ = _densify_step(raw, n2, v2, torch.Generator().manual_seed(77 + rank))
    ok &= not dp.replicas_identical(list(bad_raw.values()))
    # the bucket follows the new parameter objects
    params2 = [torch.nn.Parameter(v) for v in new_raw.values()]
    bucket.rebuild(params2)
    bucket.attach()
    ok &= bucket.flat.numel() >= sum(p.numel() for p in params2) and all(p.grad is not None for p in params2)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_parameter_bucket_and_densify_lockstep_world2():
    """The flow INTEGRATION.md §4 documents for the reference's GaussianModel: raw-parameter
    gradients are what is reduced (not the rasterizer-input gradients), and replicas stay
    bit-identical across a densify step."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_raw_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_grad_sink_set_through_the_package_reaches_the_backend():
    """`binocular3dgs_b200._C` is the compiled host side when it is built; state set on it must
    land on the ctypes backend that serves the sink path."""
    import binocular3dgs_b200 as b3
    from binocular3dgs_b200 import _backend
    from binocular3dgs_b200.dp import GradientBucket
    bucket = GradientBucket(8, 4, "cpu")
    try:
        b3._C.grad_sink = bucket
        assert _backend.native().grad_sink is bucket
        b3._C.in_autograd = True
        assert _backend.native().in_autograd is True
    finally:
        b3._C.grad_sink = None
        b3._C.in_autograd = False
    assert _backend.native().grad_sink is None
    # the sink protocol: first backward of a step overwrites; later ones are fresh tensors under
    # autograd (which adds), kernel-side accumulation without it; all_reduce ends the step
    v, acc = bucket.acquire(autograd=True)
    assert v is bucket.views() and acc is False
    assert bucket.acquire(autograd=True) == (None, False)
    v, acc = bucket.acquire(autograd=False)
    assert v is bucket.views() and acc is True
    bucket.all_reduce()
    assert bucket.acquire(autograd=False) == (bucket.views(), False)
