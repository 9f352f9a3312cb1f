#!/usr/bin/env python
"""bench.py — training views/sec (forward + backward rasterization) on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic input.  The default
workload is BASELINE.json configs[3], the largest single-GPU configuration: DTU-sized,
1 000 000 Gaussians at 1600x1200, one view (= one rasterizer forward + one backward) per
step.  configs[1] (lego, 200k at 800x800) and configs[2]/[4] (LLFF fern binocular pair,
300k at 1008x756, two renders per step, depth gradient on the first only —
train.py:100,128,149) are measured in the same run and reported under "configs".

What is timed
  value   whole-job views/s with every input already resident in HBM, through the
          `_C.rasterize_gaussians` / `_C.rasterize_gaussians_backward` pair (the C-ABI under
          it): K steps, each bracketed by CUDA events on the launch stream, an L2 flush
          (512 MiB memset) between steps outside the event pairs, barrier + synchronize on
          both sides of the region, MAX over ranks.  For N>1 every rank renders its own
          camera of the replicated scene and the per-Gaussian gradients are summed inside
          the step by b3gs_peer_allreduce over NVLink peer memory (NCCL fallback).
          scaling = weak.
  e2e     the same metric through the public API a user calls (GaussianRasterizer +
          autograd), with HOST buffers: each step copies that view's camera matrices and
          8-bit ground-truth image host->device from pinned memory, renders, takes an L1 loss,
          backpropagates to every Gaussian attribute and reads the loss back to the host.
  roofline  the dominant kernel (by measured device time) against the HBM peak in
          MEASURED_PEAKS.json; algorithmic bytes from SURVEY.md §8(d) / DESIGN.md §5.
  cpu_baseline  the CPU oracle (oracle/liboracle.so, all host threads) on a bounded
          sample of the same workload, rank 0 at N=1 only.

--impl reference runs the UNMODIFIED reference: the stock `diff_gaussian_rasterization`
package built by its own setup.py into baseline/_ref (baseline/build_reference.sh), i.e.
its own __init__.py -> _C (ext.cpp) -> rasterize_points.cu -> cuda_rasterizer/*.cu on the
same GPU, same workload, same loops.  That process never imports binocular3dgs_b200 and
maps none of this repository's shared libraries (asserted, and reported in the line).  The
reference has no CPU implementation of this path; BASELINE.json's target is ">= 2x the
reference diff-gaussian-rasterization on 1x B200".  With N>1 the reference arm runs N
independent replicas (its own multi-GPU story, script/run_llff.py) with no collective.

--impl veneer (not run by the driver) is round 1's second arm: the reference's kernels
compiled behind this repository's C-ABI (oracle/_ref/libdgr_ref.so) under this repository's
host code — it isolates kernels from host code, and carries the reference-kernel numbers
of the §8(f) rows.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from workloads import CONFIGS, make_camera, make_pixel_grads, make_scene  # noqa: E402  (pure torch/numpy)

METRIC = "training views/sec (fwd+bwd raster)"
UNIT = "views/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


# --------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def tile_work(n_contrib_hw, W, H):
    """I_f = sum over tiles of L_t = max n_contrib in the tile."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pad = torch.zeros(gy * 16, gx * 16, dtype=torch.int64, device=n_contrib_hw.device)
    pad[:H, :W] = n_contrib_hw.view(H, W).long()
    return int(pad.view(gy, 16, gx, 16).amax(dim=(1, 3)).sum())


def algorithmic_bytes(P, R, T, N, D, I_f):
    """Algorithmic bytes per launch of each stage (adapted in DESIGN.md §5
    where this library's algorithm moves fewer bytes than the reference's)."""
    return {
        "preprocess": P * (44 + 12 * (D + 1) ** 2) + 75 * P,
        # 4 passes x (4 B hist read + 8 B read + 8 B write) + gather-scan of tiles_touched
        # 4 passes x (4 B hist read + 8 B read + 8 B write)
        "depth_sort": P * 4 * 20,
        # direct tile binning (binning.cu): 4 B per instance; per (batch, tile) cell 4 B count out,
        # 4 in (chunk sums), 8 (apply), 4 in (scatter) + 8 one-byte per-warp counts out and in;
        # 12 B of rectangle + id per Gaussian, twice; 8 B per tile of ranges
        "binning": 4 * R + 36 * (-(-P // (1024 if P >= (1 << 19) else 512))) * T + 24 * P + 8 * T,
        "composite_forward": 44 * I_f + 8 * T + 24 * N,
        "grad_zero": 48 * P,
        "composite_backward": 44 * I_f + 8 * T + 28 * N + 80 * I_f,
        "preprocess_backward": P * (111 + 12 * (D + 1) ** 2) + P * (40 + 12 * (D + 1) ** 2),
    }


def mapped_repo_libraries():
    """Shared objects of THIS repository mapped into the process (/proc/self/maps)."""
    libs = set()
    try:
        for line in open("/proc/self/maps"):
            p = line.split()[-1]
            if p.endswith(".so") and p.startswith(ROOT + os.sep):
                libs.add(os.path.relpath(p, ROOT))
    except OSError:
        pass
    return sorted(libs)


# --------------------------------------------------------------------------- the arms
class Arm:
    """What a measurement loop needs from an implementation: the `_C`-level pair and the
    public surface.  `C` exposes rasterize_gaussians / rasterize_gaussians_backward with
    the reference's signatures (rasterize_points.h:18-65)."""
    name = "?"
    is_native = False
    null_grads = False          # may dL/ddepth and dL/dalpha travel as None?
    C = None
    GaussianRasterizer = None
    GaussianRasterizationSettings = None
    describe = ""


def load_native():
    import binocular3dgs_b200 as pkg
    from binocular3dgs_b200 import _backend
    arm = Arm()
    arm.name, arm.is_native, arm.null_grads = "native", True, True
    # the device-timed loop drives the C-ABI through the ctypes host side (per-stage profiling,
    # the DP gradient sink); the public surface (e2e) runs on the compiled host side when built
    arm.C = _backend.native()
    arm.GaussianRasterizer = pkg.GaussianRasterizer
    arm.GaussianRasterizationSettings = pkg.GaussianRasterizationSettings
    arm.describe = arm.C.version() + "; public surface on " + getattr(pkg._C, "name", "?")
    return arm


def load_reference():
    """The stock reference package from baseline/_ref, and nothing of this repository."""
    pkg_dir = os.path.join(REF_DIR, "diff_gaussian_rasterization")
    so = [f for f in (os.listdir(pkg_dir) if os.path.isdir(pkg_dir) else []) if f.startswith("_C") and f.endswith(".so")]
    if not so:
        return None
    sys.path.insert(0, REF_DIR)           # ahead of ROOT, whose diff_gaussian_rasterization/ is OUR drop-in
    import diff_gaussian_rasterization as dgr
    assert os.path.abspath(dgr.__file__).startswith(REF_DIR + os.sep), dgr.__file__
    arm = Arm()
    arm.name = "reference"
    arm.C = dgr._C
    arm.GaussianRasterizer = dgr.GaussianRasterizer
    arm.GaussianRasterizationSettings = dgr.GaussianRasterizationSettings
    info = {}
    try:
        info = json.load(open(os.path.join(REF_DIR, "BUILD_INFO.json")))
    except Exception:
        pass
    arm.describe = "stock diff_gaussian_rasterization (%s), built by its own setup.py, %s" % (
        info.get("extension", so[0]), info.get("arch", "sm_100a"))
    return arm


def load_veneer():
    from oracle import refbackend
    if not refbackend.available():
        return None
    from binocular3dgs_b200.rasterizer import GaussianRasterizationSettings, make_surface
    arm = Arm()
    arm.name = "veneer"
    arm.C = refbackend.reference()
    S = make_surface(arm.C)
    arm.GaussianRasterizer, arm.GaussianRasterizationSettings = S.GaussianRasterizer, GaussianRasterizationSettings
    arm.describe = "reference kernels (oracle/_ref/libdgr_ref.so) under this repository's host code"
    return arm


# --------------------------------------------------------------------------- workloads
class Workload:
    """One BASELINE.json configuration as seeded synthetic input, resident on `dev`."""

    def __init__(self, name, kind, dev, rank, pair=False):
        cfg = CONFIGS[name]
        self.name, self.kind, self.pair, self.dev = name, kind, pair, dev
        self.W, self.H, self.P = cfg["width"], cfg["height"], cfg["P"]
        self.scene_c = make_scene(self.P, seed=0, kind=kind)
        self.scene = self.scene_c.to(dev)
        self.M = self.scene.shs.shape[1]
        self.n_views = 8
        # rank-distinct cameras: a ring of views around the scene; each step cycles through 8.
        # pair: the second render of a step is the same view shifted along the camera's x axis
        # (scene/__init__.py:96-115, train.py:122-127), shift drawn once per view from +-U(0.1, 0.4)
        rs = np.random.RandomState(1234 + rank)
        self.cams_c, self.cams2_c = [], []
        for v in range(self.n_views):
            az, el = 0.3 + 0.785 * ((v + rank * 3) % 8), 0.2 - 0.05 * (v % 3)
            self.cams_c.append(make_camera(self.W, self.H, cfg["fovx"], azimuth=az, elevation=el))
            if pair:
                s = float(rs.uniform(0.1, 0.4)) * (1.0 if rs.rand() < 0.5 else -1.0)
                self.cams2_c.append(make_camera(self.W, self.H, cfg["fovx"], azimuth=az, elevation=el, shift_x=s))
        self.cams = [c.to(dev) for c in self.cams_c]
        self.cams2 = [c.to(dev) for c in self.cams2_c]
        self.bg = torch.zeros(3, device=dev)
        self.gc, self.gd, self.ga = (t.to(dev) for t in make_pixel_grads(self.W, self.H, seed=1))
        self.zero = torch.zeros_like(self.gd)       # the reference's autograd materialises zero grads
        self.views_per_step = 2 if pair else 1

    def label(self):
        s = self.scene
        return "%s%s: %d Gaussians (%s, seed 0, SH degree %d, M=%d), %dx%d, 8 cameras on a ring%s" % (
            self.name, "_pair" if self.pair else "", self.P, self.kind, s.sh_degree, self.M, self.W, self.H,
            "; 2 renders per step (view + x-shifted view), depth gradient on the first only" if self.pair else "")


def raw_forward(C, wl, cam):
    s, e = wl.scene, torch.empty(0)
    return C.rasterize_gaussians(wl.bg, s.means3D, e, s.opacities, s.scales, s.rotations, 1.0, e,
                                 cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy,
                                 wl.H, wl.W, s.shs, s.sh_degree, cam.camera_center, False, False)


def raw_backward(C, wl, cam, out, gd, ga):
    s, e = wl.scene, torch.empty(0)
    R, color, depth, alpha, radii, geom, binning, img = out
    # as this library's autograd surface calls it: gradients autograd would discard (of the empty
    # colors_precomp / cov3D_precomp placeholders) are not materialised; the reference fills them
    kw = {"skip_unobservable": True} if getattr(C, "supports_skip_unobservable", False) else {}
    return C.rasterize_gaussians_backward(wl.bg, s.means3D, radii, e, s.scales, s.rotations, 1.0, e,
                                          cam.world_view_transform, cam.full_proj_transform, cam.tanfovx,
                                          cam.tanfovy, wl.gc, gd, ga, s.shs, s.sh_degree, cam.camera_center,
                                          geom, R, binning, img, alpha, False, **kw)


def make_step(arm, wl, bucket):
    """Device-resident step: forward(s) then backward(s) of one view (or binocular pair),
    then the gradient exchange when running data-parallel."""
    C = arm.C
    none_or_zero = None if arm.null_grads else wl.zero

    def step(i):
        v = i % wl.n_views
        out = raw_forward(C, wl, wl.cams[v])
        if wl.pair:
            out2 = raw_forward(C, wl, wl.cams2[v])
            raw_backward(C, wl, wl.cams2[v], out2, none_or_zero, none_or_zero)     # colour loss only
        raw_backward(C, wl, wl.cams[v], out, wl.gd, wl.ga)
        if bucket is not None:
            bucket.all_reduce(average=True)
        return out
    return step


def dp_selfcheck(arm, dev, rank, world):
    """Correctness of the data-parallel exchange, executed by every multi-GPU bench run (the
    driver's pytest box has one GPU): (a) b3gs_peer_allreduce against NCCL on random data,
    (b) replicas bit-identical after the exchange, (c) gradients written by the backward
    kernel straight into the bucket (grad sink) + exchange == the mean of the per-rank
    gradients gathered with NCCL, (d) replicas stay bit-identical across a densification step
    (lockstep CUDA generator + reduced statistics)."""
    from binocular3dgs_b200 import dp
    out = {}
    P, M = 10007, 4
    bucket, kind = dp.make_bucket(P, M, dev)
    out["bucket"] = kind
    g = torch.Generator(device="cpu").manual_seed(4321 + rank)
    bucket.flat.copy_(torch.randn(bucket.flat.numel(), generator=g).to(dev))
    ref = bucket.flat.clone()
    dist.all_reduce(ref)
    ref /= world
    bucket.all_reduce(average=True)
    torch.cuda.synchronize()
    out["peer_vs_nccl_max_rel"] = float((bucket.flat - ref).abs().max() / ref.abs().max())
    out["replicas_identical_after_exchange"] = bool(dp.replicas_identical([bucket.flat]))
    # (c) the real path at a small size: rank-distinct cameras, sink vs gathered
    scene = make_scene(P, seed=5).to(dev)
    cam = make_camera(160, 120, azimuth=0.4 + 0.7 * rank).to(dev)
    wl = type("W", (), {})()
    wl.scene, wl.bg, wl.H, wl.W = scene, torch.zeros(3, device=dev), 120, 160
    wl.gc, wl.gd, wl.ga = (t.to(dev) for t in make_pixel_grads(160, 120, seed=6))
    C = arm.C
    C.grad_sink = None
    o = raw_forward(C, wl, cam)
    gr = raw_backward(C, wl, cam, o, wl.gd, wl.ga)   # (means2D, colors, opacity, means3D, cov3D, sh, scales, rot)
    local = dict(means3D=gr[3], shs=gr[5], opacities=gr[2], scales=gr[6], rotations=gr[7])
    gathered = {}
    for k, v in local.items():
        t = v.clone()
        dist.all_reduce(t)
        gathered[k] = t / world
    C.grad_sink = bucket
    worst, identical, keep = 0.0, True, getattr(bucket, "overlap", False)
    try:
        # backward then one fused exchange, and (symmetric-memory bucket) the chunk-pipelined
        # b3gs_backward_exchange: both must give the gathered mean, replicas bit-identical
        for overlap in ((False, True) if hasattr(bucket, "overlap") else (False,)):
            if hasattr(bucket, "overlap"):
                bucket.overlap = overlap
            bucket.begin_step()
            o = raw_forward(C, wl, cam)
            raw_backward(C, wl, cam, o, wl.gd, wl.ga)
            bucket.all_reduce(average=True)
            torch.cuda.synchronize()
            for k, v in gathered.items():
                d = float((bucket.views()[k].reshape(v.shape) - v).abs().max() / v.abs().max().clamp_min(1e-30))
                worst = max(worst, d)
            identical = identical and bool(dp.replicas_identical([bucket.flat]))
    finally:
        C.grad_sink = None
        if hasattr(bucket, "overlap"):
            bucket.overlap = keep
    out["sink_exchange_vs_gathered_max_rel"] = worst
    out["replicas_identical_gradients"] = identical
    # (c') raw-parameter models: dp.ParameterBucket in symmetric memory — autograd accumulates into
    # slices of one flat buffer, one fused exchange follows
    gp = torch.Generator(device="cpu").manual_seed(99)
    params = [torch.nn.Parameter(torch.randn(shape, generator=gp).to(dev)) for shape in ((P, 3), (P, 1, 3), (P, 1), (P, 4))]
    pb = dp.ParameterBucket(params, peer=True)
    pb.attach()
    gl = torch.Generator(device="cpu").manual_seed(1000 + rank)
    weights = [torch.randn(p_.shape, generator=gl).to(dev) for p_ in params]
    sum((p_ * w_).sum() + 0.5 * ((p_ * w_) ** 2).sum() for p_, w_ in zip(params, weights)).backward()
    expect = []
    for p_, w_ in zip(params, weights):
        t = (w_ + (p_.detach() * w_) * w_).clone()
        dist.all_reduce(t)
        expect.append(t / world)
    pb.all_reduce(average=True)
    torch.cuda.synchronize()
    out["parameter_bucket_vs_nccl_max_rel"] = max(float((p_.grad - e_).abs().max() / e_.abs().max()) for p_, e_ in zip(params, expect))
    out["parameter_bucket_replicas_identical"] = bool(dp.replicas_identical([p_.grad for p_ in params]))
    # (d) densification in lockstep
    dp.seed_lockstep(777)
    xyz, scaling = scene.means3D.clone(), torch.log(scene.scales)
    norm = gr[0][:, :2].norm(dim=-1, keepdim=True)                 # this rank's ||dL/dmean2D||
    visible = (o[4] > 0).float().unsqueeze(1)
    n2, v2, r2 = dp.reduce_densify_stats(norm * visible, visible, o[4])
    sel = ((n2 / v2.clamp_min(1)).squeeze(1) >= (n2 / v2.clamp_min(1)).median())
    std = torch.exp(scaling[sel])
    new_xyz = torch.cat((xyz, xyz[sel] + torch.normal(torch.zeros_like(std), std)), dim=0)   # CUDA generator
    out["densify_lockstep_identical"] = bool(dp.replicas_identical([new_xyz, r2]))
    out["densify_cloned"] = int(sel.sum())
    ok = (out["peer_vs_nccl_max_rel"] < 1e-5 and out["replicas_identical_after_exchange"]
          and worst < 1e-4 and out["replicas_identical_gradients"] and out["densify_lockstep_identical"]
          and out["parameter_bucket_vs_nccl_max_rel"] < 1e-5 and out["parameter_bucket_replicas_identical"])
    out["ok"] = bool(ok)
    return out


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "veneer"])
    ap.add_argument("--config", default="dtu", choices=sorted(CONFIGS) + ["fern_pair"])
    ap.add_argument("--kind", default="cube", choices=["cube", "shell"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other configs and the §8(f) rows")
    ap.add_argument("--cpu-sample-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line (the JSON): anything libraries print there (e.g.
    # NCCL's version banner) is sent to stderr instead.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    arm = {"native": load_native, "reference": load_reference, "veneer": load_veneer}[args.impl]()
    if arm is None:
        if rank == 0:
            why = ("baseline/_ref holds no built diff_gaussian_rasterization (run baseline/build_reference.sh)"
                   if args.impl == "reference" else "oracle/_ref/libdgr_ref.so was not built")
            os.write(json_fd, (json.dumps({"impl": args.impl, "unavailable": why}) + "\n").encode())
        return
    use_dp = world > 1 and arm.is_native

    def make_dp_bucket(w):
        if not use_dp:
            return None, None
        from binocular3dgs_b200.dp import make_bucket
        b, kind = make_bucket(w.P, w.M, dev, prefer_peer=os.environ.get("B3GS_DP", "peer") != "nccl")
        b.backwards_per_step = w.views_per_step      # the last backward of a step carries the exchange
        return b, kind

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
            flush.zero_()
        barrier()
        evs = []
        t0 = time.perf_counter()
        for i in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(i)
            b.record()
            evs.append((a, b))
            flush.zero_()          # L2 flush, outside the event pair
        barrier()
        wall = time.perf_counter() - t0
        per = [a.elapsed_time(b) for a, b in evs]
        ms = sum(per)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, per

    dp_check = dp_selfcheck(arm, dev, rank, world) if use_dp else None
    if dp_check is not None and not dp_check["ok"]:
        raise SystemExit("data-parallel self-check failed on rank %d: %r" % (rank, dp_check))

    # ---------------- the headline workload, device-resident (value)
    pair = args.config == "fern_pair"
    wl = Workload("fern" if pair else args.config, args.kind, dev, rank, pair=pair)
    W, H, P, M = wl.W, wl.H, wl.P, wl.M
    bucket, bucket_kind = make_dp_bucket(wl)
    C = arm.C
    if bucket is not None:
        C.grad_sink = bucket             # first backward of a step overwrites, the pair's second accumulates
    step = make_step(arm, wl, bucket)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    profiled = arm.is_native
    if profiled:
        C.profile_enable(True)
    for i in range(args.warmup):
        step(i)
        flush.zero_()
    if profiled:
        C.profile_read()                    # drop warm-up samples
    launches0 = C.launch_count() if arm.is_native else 0
    total_ms, wall, per_step = timed(step, args.steps, 0)
    prof = C.profile_read() if profiled else {}
    launches = (C.launch_count() - launches0) if arm.is_native else None
    if profiled:
        C.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * wl.views_per_step * 1000.0 / ms_per_step

    # ---------------- end to end through the public API with host buffers
    pin = lambda t: t.contiguous().pin_memory()
    host_views = []
    with torch.no_grad():
        for v in range(wl.n_views):
            hv = {}
            for tag, cam_d, cam_c in (("", wl.cams[v], wl.cams_c[v]),) + (
                    (("2", wl.cams2[v], wl.cams2_c[v]),) if wl.pair else ()):
                out = raw_forward(C, wl, cam_d)
                # ground truth as the dataset holds it: 8-bit RGB; converted on the device inside the step
                # as the reference's loader does on load (utils/general_utils.py:23, `/ 255.0`)
                hv["gt" + tag] = pin(((out[1] * 0.9 + 0.05).clamp(0, 1) * 255.0).round().to(torch.uint8).cpu())
                hv["view" + tag] = pin(cam_c.world_view_transform)
                hv["proj" + tag] = pin(cam_c.full_proj_transform)
                hv["center" + tag] = pin(cam_c.camera_center)
            host_views.append(hv)
    if arm.is_native:
        # Data-parallel: the backward writes into the bucket and autograd ADOPTS those views as
        # .grad (no pack copy); a second backward of the same step (the pair) is added by autograd
        # in place.  Single GPU: autograd owns the gradient tensors.
        C.grad_sink = bucket if use_dp else None
    leaves = [t.detach().clone().requires_grad_(True) for t in wl.scene.tensors()]
    h2d = sum(t.numel() * t.element_size() for t in host_views[0].values())
    d2h = 4

    # Double-buffered inputs: while step i computes, the copy stream uploads step i+1's
    # camera + ground-truth image from pinned memory.  The main stream waits for that
    # upload before step i's end event, so every step's event pair contains one complete
    # host->device copy (overlapped with compute, never hidden in the un-timed L2 flush).
    # The loss is copied device->host inside each step; the host consumes it one step later.
    copy_stream = torch.cuda.Stream(device=dev)
    loss_hosts = [torch.zeros(1).pin_memory() for _ in range(2)]
    staged = {}
    losses = []
    Settings, Rasterizer = arm.GaussianRasterizationSettings, arm.GaussianRasterizer

    # diagnostic only (never the reported configuration): B3GS_BENCH_E2E_NO_UPLOAD=1 re-uses device copies
    # of the inputs, to separate host-link contention from the rest of the end-to-end overhead at N = 8
    no_upload = os.environ.get("B3GS_BENCH_E2E_NO_UPLOAD") == "1"
    resident = {}

    def upload(i):
        hv = host_views[i % wl.n_views]
        copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            if no_upload:
                v = i % wl.n_views
                if v not in resident:
                    resident[v] = {k: t.to(dev, non_blocking=True) for k, t in hv.items()}
                bufs = resident[v]
            else:
                bufs = {k: t.to(dev, non_blocking=True) for k, t in hv.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (bufs, ev)

    def render(bufs, tag, cam, m, s, q, o, sh):
        settings = Settings(H, W, cam.tanfovx, cam.tanfovy, wl.bg, 1.0, bufs["view" + tag], bufs["proj" + tag],
                            wl.scene.sh_degree, bufs["center" + tag], False, False)
        means2D = torch.zeros_like(m, requires_grad=True)
        return Rasterizer(settings)(means3D=m, means2D=means2D, opacities=o, shs=sh, scales=s, rotations=q)

    def e2e_step(i):
        main = torch.cuda.current_stream(dev)
        if i not in staged:            # first step of a region: nothing was prefetched yet
            upload(i)
        bufs, ev = staged.pop(i)
        main.wait_event(ev)
        for t in bufs.values():
            t.record_stream(main)
        upload(i + 1)                  # runs concurrently with this step's kernels
        v = i % wl.n_views
        m, s, q, o, sh = leaves
        color, radii, depth, alpha = render(bufs, "", wl.cams_c[v], m, s, q, o, sh)
        loss = (color - bufs["gt"].to(torch.float32) / 255.0).abs().mean() + 1e-3 * depth.mean() + 1e-3 * alpha.mean()
        if wl.pair:
            color2 = render(bufs, "2", wl.cams2_c[v], m, s, q, o, sh)[0]
            loss = loss + (color2 - bufs["gt2"].to(torch.float32) / 255.0).abs().mean()
        for t in leaves:
            t.grad = None
        loss.backward()
        if use_dp:
            bucket.load(dict(means3D=m.grad, shs=sh.grad, opacities=o.grad, scales=s.grad, rotations=q.grad))  # no-op when adopted
            bucket.all_reduce(average=True)
        lh = loss_hosts[i & 1]
        if i >= 2:
            losses.append(float(lh[0]))              # result of step i-2, long since landed
        lh.copy_(loss.detach().reshape(1), non_blocking=True)
        main.wait_event(staged[i + 1][1])            # next step's upload completes inside this step
        return None

    e2e_steps = max(10, args.steps)
    e2e_warm = max(3, args.warmup)
    for i in range(e2e_warm):
        e2e_step(i)
        flush.zero_()
    staged.clear()
    torch.cuda.synchronize()
    e2e_ms, _, _ = timed(lambda i: e2e_step(i + e2e_warm), e2e_steps, 0)
    staged.clear()
    assert all(np.isfinite(losses)), "non-finite loss in the e2e loop"
    e2e_value = world * wl.views_per_step * 1000.0 * e2e_steps / e2e_ms

    # ---------------- roofline of the dominant kernel
    roofline, kernels = None, {}
    peak, peak_src = measured_peaks()
    if arm.is_native and prof:
        if bucket is not None:
            C.grad_sink = bucket
        out = step(0)
        torch.cuda.synchronize()
        R = int(out[0])
        n_contrib = C.blob_view(out[7], "image", "n_contrib", torch.int32, W * H, W, H)
        I_f = tile_work(n_contrib, W, H)
        T, N = ((W + 15) // 16) * ((H + 15) // 16), W * H
        alg = algorithmic_bytes(P=P, R=R, T=T, N=N, D=wl.scene.sh_degree, I_f=I_f)
        for k, (ms, calls) in prof.items():
            if calls and k in alg:
                dur = ms / calls
                kernels[k] = {"ms": round(dur, 4), "share": round(ms / total_ms, 4),
                              "calls_per_step": calls / args.steps, "alg_bytes": int(alg[k]),
                              "GBps": round(alg[k] / dur / 1e6, 1)}
        dom = max(kernels, key=lambda k: kernels[k]["share"])
        achieved = kernels[dom]["GBps"]
        try:  # dram bytes per launch from the committed ncu --set full capture of this workload
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl.name, {}).get(dom)
        except Exception:
            traffic = None
        # issue-slot roofline of the FP32-issue-bound stages: warp instructions per launch (ncu
        # smsp__inst_executed.sum of the committed capture of this workload) over what 148 SMs x 4
        # schedulers can issue in the measured time at the SM clock seen during the run
        try:
            islots = json.load(open(os.path.join(ROOT, "profiles", "issue_slots.json"))).get(wl.name, {})
        except Exception:
            islots = {}
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
        for k, v in kernels.items():
            if k in islots:
                v["warp_instructions"] = int(islots[k])
                v["issue_slot_frac"] = round(islots[k] / (v["ms"] * 1e-3 * 148 * 4 * sm_hz), 4)
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "I_f": I_f, "R": R,
                    "issue_slot_frac": {k: v["issue_slot_frac"] for k, v in kernels.items() if "issue_slot_frac" in v},
                    "note": "the composite kernels are FP32-issue bound, not HBM bound (ncu: issue slots ~80% busy, "
                            "DRAM ~1%); the HBM-bound kernels are listed with their own GB/s under `kernels`; "
                            "see DESIGN.md §5"}

    # ---------------- the other BASELINE.json configurations, device-resident, brief
    configs = None
    if not args.no_extra:
        configs = {}
        for cname, cpair in (("lego", False), ("fern", True)):
            if cname == wl.name and cpair == wl.pair:
                continue
            key = cname + ("_pair" if cpair else "")
            try:
                w2 = Workload(cname, args.kind, dev, rank, pair=cpair)
                b2, _ = make_dp_bucket(w2)
                if arm.is_native:
                    C.grad_sink = b2
                st2 = make_step(arm, w2, b2)
                n2 = max(20, min(args.steps, 50))
                ms2, _, _ = timed(st2, n2, 5)
                configs[key] = {"workload": w2.label(), "ms_per_step": round(ms2 / n2, 4), "steps": n2,
                                "value": round(world * w2.views_per_step * 1000.0 * n2 / ms2, 2), "unit": UNIT}
                del w2, b2, st2
            except Exception as ex:      # never let a secondary row break the headline line
                configs[key] = {"error": repr(ex)}
        if arm.is_native:
            C.grad_sink = None

    # ---------------- the headline step as ONE CUDA graph (the no-sync forward leaves nothing that blocks the
    # host): forward + backward captured once, replayed with the next camera written into the captured tensors
    graph_row = None
    if not args.no_extra and world == 1 and arm.is_native and getattr(C, "nosync_supported", lambda *a: False)(P, W, H):
        try:
            s_, e_ = wl.scene, torch.empty(0)
            cam0 = wl.cams[0]
            view, proj, center = (t.clone() for t in (cam0.world_view_transform, cam0.full_proj_transform,
                                                      cam0.camera_center))
            kw = {"skip_unobservable": True} if getattr(C, "supports_skip_unobservable", False) else {}

            def g_forward(capacity):
                return C.rasterize_gaussians_nosync(capacity, wl.bg, s_.means3D, e_, s_.opacities, s_.scales, s_.rotations,
                                                    1.0, e_, view, proj, cam0.tanfovx, cam0.tanfovy, wl.H, wl.W, s_.shs,
                                                    s_.sh_degree, center, False, False)

            def g_backward(out, R_):
                return C.rasterize_gaussians_backward(wl.bg, s_.means3D, out[4], e_, s_.scales, s_.rotations, 1.0, e_, view,
                                                      proj, cam0.tanfovx, cam0.tanfovy, wl.gc, wl.gd, wl.ga, s_.shs,
                                                      s_.sh_degree, center, out[5], R_, out[6], out[7], out[3], False, **kw)

            R_max = max(int(raw_forward(C, wl, c)[0]) for c in wl.cams)
            capacity = int(1.25 * R_max) + 65536
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    g_backward(g_forward(capacity), R_max)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            cuda_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cuda_graph):
                g_out = g_forward(capacity)
                g_grads = g_backward(g_out, R_max)       # R only selects the backward's kernel shape

            def replay(i):
                c = wl.cams[i % wl.n_views]
                view.copy_(c.world_view_transform); proj.copy_(c.full_proj_transform); center.copy_(c.camera_center)
                cuda_graph.replay()

            ng = max(20, min(args.steps, 50))
            ms_g, _, _ = timed(replay, ng, 5)
            # the replay must be the same computation: compare the last replayed view with the eager path
            c_last = wl.cams[(ng - 1) % wl.n_views]
            eager = raw_forward(C, wl, c_last)
            graph_row = {"what": "forward (no-sync entry, capacity 1.25 x the largest R of the 8 views) + backward of the "
                                 "headline workload captured into one CUDA graph; per step: 3 small device copies of the "
                                 "camera + one replay", "ms_per_step": round(ms_g / ng, 4), "steps": ng,
                         "value": round(wl.views_per_step * 1000.0 * ng / ms_g, 2), "unit": UNIT,
                         "image_equals_eager": bool(torch.equal(g_out[1], eager[1]))}
            del cuda_graph, g_out, g_grads
        except Exception as ex:
            graph_row = {"error": repr(ex)}

    # ---------------- config 4's growth variant: densification-like buffer growth
    # P grows by 10 % every 100 steps, as densify_and_prune replaces every parameter tensor by a
    # longer one (scene/gaussian_model.py:334-345, :393); through the public API, device-resident.
    growth = None
    if not args.no_extra and world == 1:
        try:
            P0, every, n_steps = wl.P, 100, 300
            big = make_scene(int(P0 * 1.1 ** (n_steps // every)) + 8, seed=0, kind=args.kind).to(dev)
            torch.cuda.synchronize()
            stats0 = torch.cuda.memory_stats(dev)
            sizes, evs, cur, leaves_g = [], [], 0, None
            cam0 = wl.cams[0]
            settings = Settings(H, W, cam0.tanfovx, cam0.tanfovy, wl.bg, 1.0, cam0.world_view_transform,
                                cam0.full_proj_transform, big.sh_degree, cam0.camera_center, False, False)
            for i in range(-3, n_steps):          # three untimed warm-up steps at the initial size
                Pi = int(P0 * 1.1 ** (max(i, 0) // every))
                if Pi != cur:       # "densify": every tensor is re-created at the new length
                    cur = Pi
                    leaves_g = [t[:Pi].clone().requires_grad_(True) for t in big.tensors()]
                m, sc, q, o, sh = leaves_g
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                means2D = torch.zeros_like(m, requires_grad=True)
                color, radii, depth, alpha = Rasterizer(settings)(means3D=m, means2D=means2D, opacities=o, shs=sh,
                                                                  scales=sc, rotations=q)
                for t in leaves_g:
                    t.grad = None
                torch.autograd.backward([color, depth, alpha], [wl.gc, wl.gd, wl.ga])
                b_.record()
                if i >= 0:
                    evs.append((a_, b_)); sizes.append(Pi)
            torch.cuda.synchronize()
            ms = np.array([a_.elapsed_time(b_) for a_, b_ in evs])
            stats1 = torch.cuda.memory_stats(dev)
            phases = {}
            for Pi in sorted(set(sizes)):
                sel = ms[np.array(sizes) == Pi]
                phases[str(Pi)] = {"median_ms": round(float(np.median(sel)), 4), "max_ms": round(float(sel.max()), 4),
                                   "first_ms": round(float(sel[0]), 4)}
            # a step is compared with the median of ITS size (the work itself grows with P)
            worst = max(v["max_ms"] / v["median_ms"] for v in phases.values())
            growth = {"what": "P += 10 %% every %d steps for %d steps (%d -> %d Gaussians), public API, fwd+bwd, "
                              "no L2 flush" % (every, n_steps, sizes[0], sizes[-1]),
                      "phases": phases, "worst_step_over_phase_median": round(float(worst), 3),
                      "steps_over_1p5x_median": int(sum((ms[np.array(sizes) == int(k)] > 1.5 * v["median_ms"]).sum()
                                                        for k, v in phases.items())),
                      "views_per_s": round(1000.0 * n_steps / float(ms.sum()), 2),
                      "allocator": {"cudaMalloc_calls": int(stats1.get("num_device_alloc", 0) - stats0.get("num_device_alloc", 0)),
                                    "alloc_retries": int(stats1.get("num_alloc_retries", 0) - stats0.get("num_alloc_retries", 0)),
                                    "reserved_MiB_before": int(stats0.get("reserved_bytes.all.current", 0)) >> 20,
                                    "reserved_MiB_after": int(stats1.get("reserved_bytes.all.current", 0)) >> 20,
                                    "peak_allocated_MiB": int(stats1.get("allocated_bytes.all.peak", 0)) >> 20}}
            del big, leaves_g
        except Exception as ex:
            growth = {"error": repr(ex)}

    # ---------------- CPU baseline (oracle port), rank 0 only, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and arm.is_native and not args.no_cpu_baseline:   # N=1 only (torchrun pins OMP threads to 1)
        from oracle import cpu_oracle as orc
        a = [t.numpy() for t in wl.scene_c.tensors()]
        gcn, gdn, gan = (t.cpu().numpy() for t in (wl.gc, wl.gd, wl.ga))
        views, t0 = 0, time.perf_counter()
        while True:
            cam = wl.cams_c[views % wl.n_views]
            cam_args = (cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(), cam.camera_center.numpy())
            of = orc.rasterize_forward(a[0], a[1], a[2], a[3], a[4], *cam_args, np.zeros(3, np.float32), W, H,
                                       cam.tanfovx, cam.tanfovy, wl.scene_c.sh_degree)
            orc.rasterize_backward(of, a[0], a[1], a[2], a[4], *cam_args, np.zeros(3, np.float32), W, H, cam.tanfovx,
                                   cam.tanfovy, wl.scene_c.sh_degree, gcn, gdn, gan)
            views += 1
            el = time.perf_counter() - t0
            if el >= args.cpu_sample_seconds or views >= 64:
                break
        cpu_baseline = {"value": round(views / el, 4), "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                        "sample": "%d views (fwd+bwd) of the same workload in %.1f s, OpenMP over tiles" % (views, el)}

    # ---------------- §8(f) rows next to the rasterizer
    next_rows = None
    if rank == 0 and world == 1 and not args.no_extra and args.impl == "reference":
        # the reference's training iteration from the reference's own files on the stock operator
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        try:
            next_rows = {"train_iteration": {"reference": __import__("bench_train_iteration").measure(dev, "reference")}}
        except Exception as ex:
            next_rows = {"train_iteration": {"error": repr(ex)}}
    if rank == 0 and world == 1 and not args.no_extra and args.impl in ("native", "veneer"):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        next_rows = {}

        def row(name, fn):
            try:
                next_rows[name] = fn()
            except Exception as ex:
                next_rows[name] = {"error": repr(ex)}
        if arm.is_native:
            # The native arm never executes anything under oracle/: the reference's kernels for
            # dist_cuda2 / train_iteration are timed by `--impl veneer` (same keys).
            row("photometric_loss", lambda: __import__("bench_loss").measure(dev, 3, 800, 800))
            row("binocular_loss", lambda: __import__("bench_loss").measure_binocular(dev))
            row("parameter_plumbing", lambda: __import__("bench_params").measure(dev, 200_000, 4))
            row("dist_cuda2", lambda: __import__("bench_knn").measure(1_000_000, "clustered", 3, which="native"))
            row("train_iteration", lambda: dict(
                __import__("bench_iteration").measure(dev, "fern", 20, 5, which=("native", "native_raw")),
                # the reference's own files (render, GaussianModel, losses, Adam) with only the rasterizer swapped
                dropin=__import__("bench_train_iteration").measure(dev, "dropin")))
        else:
            row("dist_cuda2", lambda: __import__("bench_knn").measure(1_000_000, "clustered", 3, which="reference"))
            row("train_iteration",
                lambda: __import__("bench_iteration").measure(dev, "fern", 20, 5, which=("reference_style",)))

    if rank == 0:
        repo_libs = mapped_repo_libraries()
        if args.impl == "reference":
            assert "binocular3dgs_b200" not in sys.modules, "the reference arm must not import the product"
            assert not [l for l in repo_libs if not l.startswith("baseline" + os.sep)], repo_libs
        per_sorted = sorted(per_step)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.label(), "P": P, "width": W, "height": H,
                       "views_per_step": wl.views_per_step,
                       "parallelism": ("view-parallel dp%d, %s of the %d-byte/Gaussian gradient bucket"
                                       % (world, "in-place two-shot all-reduce over NVLink peer memory, barriers inside "
                                          "the kernel (b3gs_peer_allreduce_fused)" if bucket_kind == "peer" else "NCCL all-reduce",
                                          4 * (11 + 3 * M))) if use_dp else
                                      ("single GPU" if world == 1 else "%d independent replicas" % world),
                       "l2": "flushed between steps (512 MiB memset outside the per-step event pairs)"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": 0 if no_upload else h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "what": "GaussianRasterizer forward + L1 loss + autograd backward; every step uploads one view's "
                            "camera + 8-bit RGB ground-truth image from pinned host memory (double-buffered on a copy stream, "
                            "converted to float on the device, completed "
                            "inside the step's event pair) and copies the loss back to pinned host memory"},
            "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": round(wall, 4),
            "step_ms": {"min": round(per_sorted[0], 4), "median": round(per_sorted[len(per_sorted) // 2], 4),
                        "max": round(per_sorted[-1], 4)},
            "impl": args.impl, "library": arm.describe, "repo_libraries_mapped": repo_libs,
        }
        if roofline:
            line["roofline"] = roofline
            line["kernels"] = kernels
        if configs:
            line["configs"] = configs
        if graph_row:
            line["graph_replay"] = graph_row
        if growth:
            line["growth"] = growth
        if dp_check:
            line["dp_check"] = dp_check
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if next_rows:
            line["next_rows"] = next_rows
        if args.impl == "reference":
            line["reference_kind"] = ("UNMODIFIED reference: baseline/_ref/diff_gaussian_rasterization (stock setup.py "
                                      "build) through its own GaussianRasterizer / _C on the same GPU; the reference "
                                      "has no CPU implementation of this path")
        elif args.impl == "veneer":
            line["reference_kind"] = arm.describe
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
