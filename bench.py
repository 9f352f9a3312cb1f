#!/usr/bin/env python
"""bench.py — training views/sec (forward + backward rasterization) on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one view = one rasterizer forward + one backward over one batch of
synthetic input (BASELINE.json configs[1]: ~200k Gaussians, 800x800, SH degree 1).

What is timed
  value   whole-job views/s with every input already resident in HBM: K steps, each
          bracketed by CUDA events on the launch stream, an L2 flush (512 MiB memset)
          between steps outside the event pairs, barrier + synchronize on both sides of
          the region, MAX over ranks.  For N>1 every rank renders its own camera of the
          replicated scene and the per-Gaussian gradients are summed by one NCCL
          all-reduce of the flat bucket the backward kernel wrote into (inside the
          step).  scaling = weak.
  e2e     the same metric through the public API a user calls (GaussianRasterizer +
          autograd), with HOST buffers: each step copies that view's camera matrices and
          ground-truth image host->device from pinned memory, renders, takes an L1 loss,
          backpropagates to every Gaussian attribute and reads the loss back to the host.
  roofline  the dominant kernel (by measured device time) against the HBM peak in
          MEASURED_PEAKS.json; algorithmic bytes from SURVEY.md §8(d).
  cpu_baseline  the CPU oracle (oracle/liboracle.so, all host threads) on a bounded
          sample of the same workload, rank 0 only.

--impl reference runs the reference's OWN CUDA kernels (oracle/_ref/libdgr_ref.so,
compiled unmodified from /root/reference) through identical host code on the same GPU:
the reference has no CPU implementation of this path, and BASELINE.json's target is
">= 2x the reference diff-gaussian-rasterization on 1x B200".  With N>1 the reference
arm runs N independent replicas (its own multi-GPU story, script/run_llff.py) with no
collective.
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

from binocular3dgs_b200.dp import GradientBucket, make_bucket  # noqa: E402
from binocular3dgs_b200.rasterizer import GaussianRasterizationSettings, make_surface  # noqa: E402
from binocular3dgs_b200.synthetic import CONFIGS, make_camera, make_pixel_grads, make_scene  # noqa: E402

METRIC = "training views/sec (fwd+bwd raster)"
UNIT = "views/s"


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


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="lego", choices=sorted(CONFIGS))
    ap.add_argument("--kind", default="cube", choices=["cube", "shell"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    if args.impl == "reference":
        from oracle import refbackend
        if not refbackend.available():
            if rank == 0:
                os.write(json_fd, (json.dumps({"impl": "reference", "unavailable":
                                               "oracle/_ref/libdgr_ref.so was not built"}) + "\n").encode())
            return
        back = refbackend.reference()
    else:
        from binocular3dgs_b200 import _backend
        back = _backend.native()
    # the public surface (e2e) runs on the compiled host side when it has been built; the
    # device-timed loop below drives the C-ABI through the ctypes host side (per-stage
    # profiling, the DP gradient sink) — same library, same kernels
    S = make_surface(_backend.preferred() if args.impl == "native" else back)

    cfg = CONFIGS[args.config]
    W, H, P = cfg["width"], cfg["height"], cfg["P"]
    scene_c = make_scene(P, seed=0, kind=args.kind)
    scene = scene_c.to(dev)
    M = scene.shs.shape[1]
    # rank-distinct cameras: a ring of views around the scene; each step cycles through 8
    n_views = 8
    cams_c = [make_camera(W, H, cfg["fovx"], azimuth=0.3 + 0.785 * ((v + rank * 3) % 8), elevation=0.2 - 0.05 * (v % 3))
              for v in range(n_views)]
    cams = [c.to(dev) for c in cams_c]
    bg = torch.zeros(3, device=dev)
    gc, gd, ga = (t.to(dev) for t in make_pixel_grads(W, H, seed=1))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    e = torch.empty(0)
    use_dp = world > 1 and args.impl == "native"
    # world > 1: the bucket lives in symmetric memory and is reduced by b3gs_peer_allreduce
    # (B3GS_DP=nccl forces the NCCL all-reduce for A/B timing)
    bucket, bucket_kind = (make_bucket(P, M, dev, prefer_peer=os.environ.get("B3GS_DP", "peer") != "nccl")
                           if args.impl == "native" else (None, None))
    if bucket is not None:
        back.grad_sink = bucket.views()

    # ---------------- device-resident step (value)
    def step(i):
        cam = cams[i % n_views]
        out = back.rasterize_gaussians(bg, scene.means3D, e, scene.opacities, scene.scales, scene.rotations, 1.0, e,
                                       cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy,
                                       H, W, scene.shs, scene.sh_degree, cam.camera_center, False, False)
        R, color, depth, alpha, radii, geom, binning, img = out
        back.rasterize_gaussians_backward(bg, scene.means3D, radii, e, scene.scales, scene.rotations, 1.0, e,
                                          cam.world_view_transform, cam.full_proj_transform, cam.tanfovx,
                                          cam.tanfovy, gc, gd, ga, scene.shs, scene.sh_degree, cam.camera_center,
                                          geom, R, binning, img, alpha, False)
        if use_dp:
            bucket.all_reduce(average=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = back.launch_count()
    back.profile_enable(True)
    back.profile_read()
    for i in range(args.warmup):
        step(i)
        flush.zero_()
    back.profile_read()                    # drop warm-up samples
    launches0 = back.launch_count()
    total_ms, wall = timed(step, args.steps, 0)
    prof = back.profile_read()
    launches = back.launch_count() - launches0
    back.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * 1000.0 / ms_per_step

    # ---------------- end to end through the public API with host buffers
    pin = lambda t: t.contiguous().pin_memory()
    host_views = []
    with torch.no_grad():
        for v in range(n_views):
            out = back.rasterize_gaussians(bg, scene.means3D, e, scene.opacities, scene.scales, scene.rotations, 1.0,
                                           e, cams[v].world_view_transform, cams[v].full_proj_transform,
                                           cams[v].tanfovx, cams[v].tanfovy, H, W, scene.shs, scene.sh_degree,
                                           cams[v].camera_center, False, False)
            gt = (out[1] * 0.9 + 0.05).clamp(0, 1).cpu()
            host_views.append(dict(gt=pin(gt), view=pin(cams_c[v].world_view_transform),
                                   proj=pin(cams_c[v].full_proj_transform), center=pin(cams_c[v].camera_center)))
    leaves = [t.detach().clone().requires_grad_(True) for t in scene.tensors()]
    loss_host = torch.zeros(1).pin_memory()
    h2d = host_views[0]["gt"].numel() * 4 + (16 + 16 + 3) * 4
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

    def upload(i):
        hv = host_views[i % n_views]
        copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            bufs = {k: hv[k].to(dev, non_blocking=True) for k in ("gt", "view", "proj", "center")}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (bufs, ev)

    def e2e_step(i):
        main = torch.cuda.current_stream(dev)
        if i not in staged:            # first step of a region: nothing was prefetched yet
            upload(i)
        bufs, ev = staged.pop(i)
        main.wait_event(ev)
        for t in bufs.values():
            t.record_stream(main)
        upload(i + 1)                  # runs concurrently with this step's kernels
        cam = cams_c[i % n_views]
        m, s, q, o, sh = leaves
        settings = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, bufs["view"], bufs["proj"],
                                                 scene.sh_degree, bufs["center"], False, False)
        means2D = torch.zeros_like(m, requires_grad=True)
        color, radii, depth, alpha = S.GaussianRasterizer(settings)(
            means3D=m, means2D=means2D, opacities=o, shs=sh, scales=s, rotations=q)
        loss = (color - bufs["gt"]).abs().mean() + 1e-3 * depth.mean() + 1e-3 * alpha.mean()
        for t in leaves:
            t.grad = None
        loss.backward()
        if use_dp:
            bucket.load(dict(means3D=m.grad, shs=sh.grad, opacities=o.grad, scales=s.grad, rotations=q.grad))
            bucket.all_reduce(average=True)
        lh = loss_hosts[i & 1]
        if i >= 2:
            losses.append(float(lh[0]))              # result of step i-2, long since landed
        lh.copy_(loss.detach().reshape(1), non_blocking=True)
        main.wait_event(staged[i + 1][1])            # next step's upload completes inside this step
        return None

    back.grad_sink = None                          # autograd owns the gradient tensors on this path
    e2e_steps = max(10, args.steps)
    e2e_warm = max(3, args.warmup)
    for i in range(e2e_warm):
        e2e_step(i)
        flush.zero_()
    staged.clear()
    torch.cuda.synchronize()
    e2e_ms, _ = timed(lambda i: e2e_step(i + e2e_warm), e2e_steps, 0)
    assert all(np.isfinite(losses)), "non-finite loss in the e2e loop"
    e2e_value = world * 1000.0 * e2e_steps / e2e_ms

    # ---------------- roofline of the dominant kernel
    roofline, kernels = None, {}
    peak, peak_src = measured_peaks()
    if args.impl == "native" and prof:
        out = step(0)
        torch.cuda.synchronize()
        R = out[0]
        n_contrib = back.blob_view(out[7], "image", "n_contrib", torch.int32, W * H, W, H)
        I_f = tile_work(n_contrib, W, H)
        T, N = ((W + 15) // 16) * ((H + 15) // 16), W * H
        D = scene.sh_degree
        alg = {  # algorithmic bytes per launch, SURVEY.md §8(d)
            "preprocess": P * (44 + 12 * (D + 1) ** 2) + 75 * P,
            # 4 passes x (4 B hist read + 8 B read + 8 B write) + gather-scan of tiles_touched
            "depth_sort": P * (4 * 20 + 16),
            # emit (tile,id) 8 B + passes x 20 B + 4 B boundary scan per instance, 8 B per tile
            "binning": 28 * P + R * (8 + 20 * (1 if T <= 256 else 2 if T <= 65536 else 3) + 4) + 8 * T,
            "composite_forward": 44 * I_f + 8 * T + 24 * N,
            "grad_zero": 48 * P,
            "composite_backward": 44 * I_f + 8 * T + 28 * N + 80 * I_f,
            "preprocess_backward": P * (111 + 12 * (D + 1) ** 2) + P * (40 + 12 * (D + 1) ** 2),
        }
        for k, (ms, calls) in prof.items():
            if calls:
                dur = ms / calls
                kernels[k] = {"ms": round(dur, 4), "share": round(ms / total_ms, 4),
                              "alg_bytes": int(alg[k]), "GBps": round(alg[k] / dur / 1e6, 1)}
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        achieved = kernels[dom]["GBps"]
        try:  # dram bytes per launch from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            traffic = None
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "I_f": I_f, "R": R,
                    "note": "composite kernels are FP32-issue bound, not HBM bound (ncu: issue slots ~83% busy, "
                            "DRAM <1%); see DESIGN.md"}

    # ---------------- CPU baseline (oracle port), rank 0 only, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # N=1 only (torchrun pins OMP threads to 1)
        from oracle import cpu_oracle as orc
        a = [t.numpy() for t in scene_c.tensors()]
        gcn, gdn, gan = (t.cpu().numpy() for t in (gc, gd, ga))
        views, t0 = 0, time.perf_counter()
        while True:
            cam = cams_c[views % n_views]
            cam_args = (cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(), cam.camera_center.numpy())
            of = orc.rasterize_forward(a[0], a[1], a[2], a[3], a[4], *cam_args, np.zeros(3, np.float32), W, H,
                                       cam.tanfovx, cam.tanfovy, scene_c.sh_degree)
            orc.rasterize_backward(of, a[0], a[1], a[2], a[4], *cam_args, np.zeros(3, np.float32), W, H, cam.tanfovx,
                                   cam.tanfovy, scene_c.sh_degree, gcn, gdn, gan)
            views += 1
            el = time.perf_counter() - t0
            if el >= args.cpu_sample_seconds or views >= 64:
                break
        cpu_baseline = {"value": round(views / el, 4), "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                        "sample": "%d views (fwd+bwd) of the same workload in %.1f s, OpenMP over tiles" % (views, el)}

    # ---------------- §8(f) rank 1: fused photometric loss, timed beside the torch-style one
    next_rows = None
    if rank == 0 and world == 1 and args.impl == "native":
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_loss
            next_rows = {"photometric_loss": bench_loss.measure(dev, 3, H, W)}
        except Exception as ex:  # never let the auxiliary row break the headline line
            next_rows = {"photometric_loss": {"error": repr(ex)}}
        try:  # §8(f) rank 2 at the size of the config that uses it (LLFF fern, 1008x756)
            next_rows["binocular_loss"] = bench_loss.measure_binocular(dev)
        except Exception as ex:
            next_rows["binocular_loss"] = {"error": repr(ex)}
        try:  # §8(f) rank 3 at the bench's own P and M
            import bench_params
            next_rows["parameter_plumbing"] = bench_params.measure(dev, P, M)
        except Exception as ex:
            next_rows["parameter_plumbing"] = {"error": repr(ex)}
        # The native arm never executes anything under oracle/: the reference's own kernels
        # for these two rows are timed by `bench.py --impl reference` (same keys, below).
        try:  # §8(f) rank 4 on an SfM-like cloud
            import bench_knn
            next_rows["dist_cuda2"] = bench_knn.measure(1_000_000, "clustered", 3, which="native")
        except Exception as ex:
            next_rows["dist_cuda2"] = {"error": repr(ex)}
        try:  # everything together: the reference's training iteration on the binocular config
            import bench_iteration
            next_rows["train_iteration"] = bench_iteration.measure(dev, "fern", 20, 5, which=("native",))
        except Exception as ex:
            next_rows["train_iteration"] = {"error": repr(ex)}
    elif rank == 0 and world == 1 and args.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        next_rows = {}
        try:  # the reference's simple-knn kernel (oracle/_ref/libknn_ref.so) on the same cloud
            import bench_knn
            next_rows["dist_cuda2"] = bench_knn.measure(1_000_000, "clustered", 3, which="reference")
        except Exception as ex:
            next_rows["dist_cuda2"] = {"error": repr(ex)}
        try:  # the reference's kernels + everything around them as the reference writes it
            import bench_iteration
            next_rows["train_iteration"] = bench_iteration.measure(dev, "fern", 20, 5, which=("reference_style",))
        except Exception as ex:
            next_rows["train_iteration"] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %d Gaussians (%s, seed 0, SH degree %d, M=%d), %dx%d, 8 cameras on a ring"
                                   % (args.config, P, args.kind, scene.sh_degree, M, W, H),
                       "P": P, "width": W, "height": H,
                       "parallelism": ("view-parallel dp%d, %s of the %d-byte/Gaussian gradient bucket"
                                       % (world, "in-place two-shot all-reduce over NVLink peer memory "
                                          "(b3gs_peer_allreduce)" if bucket_kind == "peer" else "NCCL all-reduce",
                                          4 * (11 + 3 * M))) if use_dp else
                                      ("single GPU" if world == 1 else "%d independent replicas" % world),
                       "l2": "flushed between steps (512 MiB memset outside the per-step event pairs)"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "what": "GaussianRasterizer forward + L1 loss + autograd backward; every step uploads one view's "
                            "camera + GT image from pinned host memory (double-buffered on a copy stream, completed "
                            "inside the step's event pair) and copies the loss back to pinned host memory"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": round(wall, 4),
            "impl": args.impl, "library": back.version(),
        }
        if roofline:
            line["roofline"] = roofline
            line["kernels"] = kernels
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if next_rows:
            line["next_rows"] = next_rows
        if args.impl == "reference":
            line["reference_kind"] = ("reference CUDA kernels (oracle/_ref, unmodified sources) on the same GPU; the "
                                      "reference has no CPU implementation of this path")
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
