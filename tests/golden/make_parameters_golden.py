"""Golden vectors for the per-step parameter plumbing, produced by running the reference's
own statements (scene/gaussian_model.py, train.py — cited at each block) with torch on CPU:

    python tests/golden/make_parameters_golden.py      # writes tests/golden/params_*.npz

GaussianModel itself cannot be imported without its CUDA extensions (simple_knn._C at module
scope, device="cuda" literals), so the statements are restated here one for one; the
arithmetic under test is torch's (sigmoid/exp/normalize/cat and torch.optim.Adam).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"params_sh1": dict(P=257, M=4, seed=21, steps=4), "params_sh3": dict(P=130, M=16, seed=22, steps=3),
         "params_sh0": dict(P=64, M=1, seed=23, steps=2)}
# arguments/__init__.py:75-82 (position_lr_init * spatial_lr_scale with spatial_lr_scale = 1)
LRS = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20.0, opacity=0.05, scaling=0.005, rotation=0.001)
ORDER = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def make_raw(P, M, seed, **_):
    g = torch.Generator().manual_seed(seed)
    raw = dict(xyz=torch.randn(P, 3, generator=g), f_dc=torch.randn(P, 1, 3, generator=g),
               f_rest=0.1 * torch.randn(P, M - 1, 3, generator=g), opacity=2.0 * torch.randn(P, 1, generator=g),
               scaling=torch.randn(P, 3, generator=g) - 3.0, rotation=torch.randn(P, 4, generator=g))
    raw["rotation"][0] = 0.0                      # normalize's eps branch
    raw["rotation"][1] = torch.tensor([2.0, 0.0, 0.0, 0.0])
    return raw


def make_upstream(P, M, seed, step=0):
    g = torch.Generator().manual_seed(seed * 1000 + step)
    return dict(shs=torch.randn(P, M, 3, generator=g), opacities=torch.randn(P, 1, generator=g),
                scales=torch.randn(P, 3, generator=g), rotations=torch.randn(P, 4, generator=g),
                xyz=1e-3 * torch.randn(P, 3, generator=g))


def make_view_stats(P, seed, step=0):
    g = torch.Generator().manual_seed(seed * 77 + step)
    radii = torch.randint(-1, 40, (P,), generator=g, dtype=torch.int32).clamp_min(0)
    return radii, 1e-3 * torch.randn(P, 3, generator=g)


def reference_activations(raw):
    """scene/gaussian_model.py:95-115"""
    features = torch.cat((raw["f_dc"], raw["f_rest"]), dim=1)
    return features, torch.sigmoid(raw["opacity"]), torch.exp(raw["scaling"]), torch.nn.functional.normalize(raw["rotation"])


def main():
    for name, c in CASES.items():
        P, M = c["P"], c["M"]
        params = {k: torch.nn.Parameter(v.clone()) for k, v in make_raw(**c).items()}
        out = {}
        # ---- activations forward + autograd backward
        up = make_upstream(P, M, c["seed"])
        acts = reference_activations(params)
        torch.autograd.backward(list(acts), [up["shs"], up["opacities"], up["scales"], up["rotations"]])
        for k, a in zip(("shs", "opacities", "scales", "rotations"), acts):
            out["act_" + k] = a.detach().numpy()
        for k in ORDER[1:]:
            out["actgrad_" + k] = params[k].grad.numpy().copy()
        # ---- Adam as gaussian_model.py:154-163 builds it, a few steps with fresh gradients
        groups = [{"params": [params[k]], "lr": LRS[k], "name": k} for k in ORDER]
        opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        for step in range(c["steps"]):
            up = make_upstream(P, M, c["seed"], step + 1)
            opt.zero_grad(set_to_none=True)
            acts = reference_activations(params)
            loss = sum((a * up[k]).sum() for a, k in zip(acts, ("shs", "opacities", "scales", "rotations")))
            loss = loss + (params["xyz"] * up["xyz"]).sum()
            loss.backward()
            opt.step()
        for k in ORDER:
            st = opt.state[params[k]]
            out["adam_" + k] = params[k].detach().numpy().copy()
            out["adam_m_" + k] = st["exp_avg"].numpy().copy()
            out["adam_v_" + k] = st["exp_avg_sq"].numpy().copy()
        # ---- opacity decay, gaussian_model.py:307-309 with utils/general_utils.py:18-19
        raw = make_raw(**c)
        opacity = torch.sigmoid(raw["opacity"]) * 0.995
        out["decay_0995"] = torch.log(opacity / (1 - opacity)).numpy()
        # ---- densification statistics over three views, train.py:170-171, gaussian_model.py:409-411
        accum, denom, max_radii2D = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)
        for view in range(3):
            radii, vgrad = make_view_stats(P, c["seed"], view)
            visibility_filter = radii > 0
            max_radii2D[visibility_filter] = torch.max(max_radii2D[visibility_filter], radii[visibility_filter])
            accum[visibility_filter] += torch.norm(vgrad[visibility_filter, :2], dim=-1, keepdim=True)
            denom[visibility_filter] += 1
        out.update(stats_accum=accum.numpy(), stats_denom=denom.numpy(), stats_max_radii2D=max_radii2D.numpy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "ok", {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()
