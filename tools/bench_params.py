"""Timing of the per-step parameter plumbing against the same work
written the way the reference writes it (scene/gaussian_model.py: torch activations,
torch.optim.Adam over six groups, boolean-mask densification statistics), both on the GPU,
CUDA events, L2 flushed between iterations.  Used by bench.py ("next_rows")."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LRS = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20.0, opacity=0.05, scaling=0.005, rotation=0.001)
ORDER = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def _raw(P, M, dev):
    g = torch.Generator().manual_seed(0)
    return dict(xyz=torch.randn(P, 3, generator=g).to(dev), f_dc=torch.randn(P, 1, 3, generator=g).to(dev),
                f_rest=(0.1 * torch.randn(P, M - 1, 3, generator=g)).to(dev),
                opacity=torch.randn(P, 1, generator=g).to(dev), scaling=(torch.randn(P, 3, generator=g) - 3).to(dev),
                rotation=torch.randn(P, 4, generator=g).to(dev))


def _median_ms(fn, flush, iters, warmup, setup=None):
    ts = []
    for i in range(warmup + iters):
        if setup:
            setup()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def measure(dev, P=200_000, M=4, iters=30, warmup=5):
    from binocular3dgs_b200 import parameters
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"what": "P=%d, M=%d: activations fwd+bwd, Adam step over 6 groups, densification statistics" % (P, M)}
    n_param = P * (11 + 3 * M)

    # ---- activations forward + backward
    raw = {k: v.requires_grad_(True) for k, v in _raw(P, M, dev).items()}
    ups = [torch.randn(P, M, 3, device=dev), torch.randn(P, 1, device=dev), torch.randn(P, 3, device=dev),
           torch.randn(P, 4, device=dev)]

    def torch_act():
        acts = (torch.cat((raw["f_dc"], raw["f_rest"]), dim=1), torch.sigmoid(raw["opacity"]),
                torch.exp(raw["scaling"]), torch.nn.functional.normalize(raw["rotation"]))
        torch.autograd.backward(list(acts), ups)

    def fused_act():
        acts = parameters.activate(raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"])
        torch.autograd.backward(list(acts), ups)

    def clear():
        for v in raw.values():
            v.grad = None
    out["activations"] = {"fused_ms": round(_median_ms(fused_act, flush, iters, warmup, clear), 4),
                          "torch_reference_style_ms": round(_median_ms(torch_act, flush, iters, warmup, clear), 4),
                          "alg_bytes": 4 * 4 * P * (8 + 3 * M)}

    # ---- Adam
    def build(cls):
        params = {k: torch.nn.Parameter(v) for k, v in _raw(P, M, dev).items()}
        for v in params.values():
            v.grad = torch.randn_like(v)
        return cls([{"params": [params[k]], "lr": LRS[k], "name": k} for k in ORDER], lr=0.0, eps=1e-15)
    oa, ob = build(torch.optim.Adam), build(parameters.FusedAdam)
    out["adam"] = {"fused_ms": round(_median_ms(ob.step, flush, iters, warmup), 4),
                   "torch_reference_style_ms": round(_median_ms(oa.step, flush, iters, warmup), 4),
                   "alg_bytes": 40 * n_param}
    out["adam"]["fused_GBps"] = round(out["adam"]["alg_bytes"] / out["adam"]["fused_ms"] / 1e6, 1)

    # ---- densification statistics
    radii = torch.randint(0, 30, (P,), device=dev, dtype=torch.int32)
    radii[::3] = 0
    vgrad = torch.randn(P, 3, device=dev)
    acc, den, mx = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)

    def torch_stats():
        vis = radii > 0
        mx[vis] = torch.max(mx[vis], radii[vis])
        acc[vis] += torch.norm(vgrad[vis, :2], dim=-1, keepdim=True)
        den[vis] += 1

    def fused_stats():
        parameters.add_densification_stats(vgrad, radii, acc, den, mx)
    out["densify_stats"] = {"fused_ms": round(_median_ms(fused_stats, flush, iters, warmup), 4),
                            "torch_reference_style_ms": round(_median_ms(torch_stats, flush, iters, warmup), 4)}
    f = sum(out[k]["fused_ms"] for k in ("activations", "adam", "densify_stats"))
    t = sum(out[k]["torch_reference_style_ms"] for k in ("activations", "adam", "densify_stats"))
    out["fused_ms"], out["torch_reference_style_ms"], out["speedup"] = round(f, 4), round(t, 4), round(t / f, 2)
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(measure(torch.device("cuda:0"))))
