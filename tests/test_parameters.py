"""Per-step parameter plumbing.

CPU: the numpy oracle (oracle/parameters_oracle.py) against golden vectors produced by the
reference's statements run with torch on CPU (tests/golden/make_parameters_golden.py).
GPU (-m gpu): the CUDA kernels through the C-ABI against the same golden vectors, at full
size against the oracle, and FusedAdam against torch.optim.Adam on the same device through
the reference's optimizer surgery (state concatenation as in cat_tensors_to_optimizer).

Tolerances (float32 arithmetic against float64 / against torch's float32): 2e-6 relative
to the tensor's max for activations, gradients and Adam moments; Adam-updated parameters
to 1e-6 absolute + 2e-6 relative (with eps = 1e-15 the update m/sqrt(v) is a ratio of two
rounded quantities); densification statistics 1e-6 relative, counters and maxima exact."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import parameters_oracle as po

TOL = 2e-6
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_parameters_golden",
                                              os.path.join(HERE, "golden", "make_parameters_golden.py"))
mpg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mpg)
NAMES = sorted(mpg.CASES)
ACT = ("shs", "opacities", "scales", "rotations")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape or a.size == b.size
    if b.size == 0:
        return 0.0
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-30))


def close_abs_rel(a, b, atol=1e-6, rtol=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return b.size == 0 or np.abs(a - b).max() < atol + rtol * np.abs(b).max()


def load(name):
    c = mpg.CASES[name]
    return c, np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_torch_golden(name):
    c, g = load(name)
    raw = {k: v.numpy() for k, v in mpg.make_raw(**c).items()}
    acts = po.activate(raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"])
    for k, a in zip(ACT, acts):
        assert rel(a, g["act_" + k]) < TOL, k
    up = {k: v.numpy() for k, v in mpg.make_upstream(c["P"], c["M"], c["seed"]).items()}
    grads = po.activate_grad(raw["opacity"], raw["scaling"], raw["rotation"], up["shs"], up["opacities"], up["scales"],
                             up["rotations"])
    for k, a in zip(mpg.ORDER[1:], grads):
        assert rel(a, g["actgrad_" + k]) < TOL, k
    # Adam: replay the golden run's gradients through the oracle's update
    p = {k: raw[k].astype(np.float64) for k in mpg.ORDER}
    m = {k: np.zeros_like(p[k]) for k in mpg.ORDER}
    v = {k: np.zeros_like(p[k]) for k in mpg.ORDER}
    for step in range(c["steps"]):
        up = {k: t.numpy() for k, t in mpg.make_upstream(c["P"], c["M"], c["seed"], step + 1).items()}
        gr = dict(zip(mpg.ORDER[1:], po.activate_grad(p["opacity"], p["scaling"], p["rotation"], up["shs"],
                                                      up["opacities"], up["scales"], up["rotations"])))
        gr["xyz"] = up["xyz"]
        for k in mpg.ORDER:
            p[k], m[k], v[k] = po.adam_step(p[k], gr[k], m[k], v[k], step + 1, mpg.LRS[k])
    for k in mpg.ORDER:
        assert close_abs_rel(p[k], g["adam_" + k]), k
        assert rel(m[k], g["adam_m_" + k]) < 1e-5 and rel(v[k], g["adam_v_" + k]) < 1e-5, k
    assert rel(po.opacity_decay(raw["opacity"], 0.995), g["decay_0995"]) < TOL
    accum, denom, mx = np.zeros(c["P"]), np.zeros(c["P"]), np.zeros(c["P"])
    for view in range(3):
        radii, vgrad = mpg.make_view_stats(c["P"], c["seed"], view)
        accum, denom, mx = po.densify_stats(vgrad.numpy(), radii.numpy(), accum, denom, mx)
    assert rel(accum, g["stats_accum"][:, 0]) < 1e-6
    assert (denom == g["stats_denom"][:, 0]).all() and (mx == g["stats_max_radii2D"]).all()


def test_api_surface_and_loud_failures():
    from binocular3dgs_b200 import parameters
    assert issubclass(parameters.FusedAdam, torch.optim.Adam)
    p = torch.nn.Parameter(torch.zeros(4, 3))
    opt = parameters.FusedAdam([{"params": [p], "lr": 0.1, "name": "xyz"}], lr=0.0, eps=1e-15)
    assert opt.param_groups[0]["name"] == "xyz" and opt.param_groups[0]["eps"] == 1e-15
    p.grad = torch.ones(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()
    with pytest.raises(NotImplementedError):
        parameters.FusedAdam([p], amsgrad=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        parameters.activate(torch.zeros(2, 1, 3), torch.zeros(2, 3, 3), torch.zeros(2, 1), torch.zeros(2, 3),
                            torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        parameters.opacity_decay(torch.zeros(2, 1))


# ------------------------------------------------------------------------------- GPU
def _cuda_params(c):
    return {k: torch.nn.Parameter(v.cuda()) for k, v in mpg.make_raw(**c).items()}


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_kernels_match_torch_golden(name):
    from binocular3dgs_b200 import parameters
    c, g = load(name)
    P, M = c["P"], c["M"]
    params = _cuda_params(c)
    up = {k: v.cuda() for k, v in mpg.make_upstream(P, M, c["seed"]).items()}
    acts = parameters.activate(params["f_dc"], params["f_rest"], params["opacity"], params["scaling"], params["rotation"])
    torch.autograd.backward(list(acts), [up[k] for k in ACT])
    for k, a in zip(ACT, acts):
        assert rel(a.detach().cpu().numpy(), g["act_" + k]) < TOL, k
    for k in mpg.ORDER[1:]:
        assert rel(params[k].grad.cpu().numpy(), g["actgrad_" + k]) < TOL, k
    groups = [{"params": [params[k]], "lr": mpg.LRS[k], "name": k} for k in mpg.ORDER]
    opt = parameters.FusedAdam(groups, lr=0.0, eps=1e-15)
    for step in range(c["steps"]):
        up = {k: v.cuda() for k, v in mpg.make_upstream(P, M, c["seed"], step + 1).items()}
        opt.zero_grad(set_to_none=True)
        acts = parameters.activate(params["f_dc"], params["f_rest"], params["opacity"], params["scaling"],
                                   params["rotation"])
        loss = sum((a * up[k]).sum() for a, k in zip(acts, ACT)) + (params["xyz"] * up["xyz"]).sum()
        loss.backward()
        opt.step()
    for k in mpg.ORDER:
        st = opt.state[params[k]]
        assert float(st["step"]) == c["steps"]
        want = g["adam_" + k]
        assert close_abs_rel(params[k].detach().cpu().numpy(), want), k
        assert rel(st["exp_avg"].cpu().numpy(), g["adam_m_" + k]) < 1e-5, k
        assert rel(st["exp_avg_sq"].cpu().numpy(), g["adam_v_" + k]) < 1e-5, k
    raw = mpg.make_raw(**c)
    op = raw["opacity"].cuda()
    assert parameters.opacity_decay(op, 0.995) is op
    assert rel(op.cpu().numpy(), g["decay_0995"]) < TOL
    accum, denom, mx = (torch.zeros(P, 1).cuda(), torch.zeros(P, 1).cuda(), torch.zeros(P).cuda())
    for view in range(3):
        radii, vgrad = mpg.make_view_stats(P, c["seed"], view)
        parameters.add_densification_stats(vgrad.cuda(), radii.cuda(), accum, denom, mx)
    assert rel(accum.cpu().numpy(), g["stats_accum"]) < 1e-6
    assert (denom.cpu().numpy() == g["stats_denom"]).all() and (mx.cpu().numpy() == g["stats_max_radii2D"]).all()


@pytest.mark.gpu
def test_fused_adam_follows_torch_adam_through_densification_surgery():
    """200k Gaussians, 6 groups, 3 steps, then the state concatenation of
    cat_tensors_to_optimizer (gaussian_model.py:311-331) and 2 more steps; torch.optim.Adam
    on the same device runs the identical sequence."""
    from binocular3dgs_b200 import parameters
    c = dict(P=200_003, M=4, seed=5)
    g = torch.Generator().manual_seed(9)

    def build(cls):
        params = _cuda_params(c)
        groups = [{"params": [params[k]], "lr": mpg.LRS[k], "name": k} for k in mpg.ORDER]
        return params, cls(groups, lr=0.0, eps=1e-15)

    (pa, oa), (pb, ob) = build(torch.optim.Adam), build(parameters.FusedAdam)

    def grow(opt, extra):
        for group in opt.param_groups:
            ext = extra[group["name"]]
            st = opt.state.get(group["params"][0], None)
            st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(ext)), dim=0)
            st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
            del opt.state[group["params"][0]]
            group["params"][0] = torch.nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
            opt.state[group["params"][0]] = st

    for step in range(5):
        if step == 3:
            extra = {k: v[:1000].detach().clone() for k, v in pa.items()}
            grow(oa, extra)
            grow(ob, {k: v.clone() for k, v in extra.items()})
        for opt in (oa, ob):
            gen = torch.Generator().manual_seed(100 + step)
            for group in opt.param_groups:
                p = group["params"][0]
                scale = 1e-3 if group["name"] == "xyz" else 1.0
                grad = (scale * torch.randn(p.shape, generator=gen)).cuda()
                grad[::7] = 0.0          # Gaussians outside the frustum receive exact zeros
                p.grad = grad
            opt.step()
    for ga, gb in zip(oa.param_groups, ob.param_groups):
        a, b = ga["params"][0], gb["params"][0]
        assert a.shape == b.shape and a.shape[0] == c["P"] + 1000
        assert float((a - b).abs().max()) < 1e-6 + TOL * float(a.abs().max()), ga["name"]
        assert rel(ob.state[b]["exp_avg"].cpu().numpy(), oa.state[a]["exp_avg"].cpu().numpy()) < 1e-6
        assert rel(ob.state[b]["exp_avg_sq"].cpu().numpy(), oa.state[a]["exp_avg_sq"].cpu().numpy()) < 1e-6


@pytest.mark.gpu
def test_activate_full_size_vs_oracle_and_partial_gradients():
    from binocular3dgs_b200 import parameters
    c = dict(P=100_001, M=16, seed=8)
    raw = mpg.make_raw(**c)
    params = {k: v.cuda().requires_grad_(k != "f_rest") for k, v in raw.items()}     # f_rest frozen
    acts = parameters.activate(params["f_dc"], params["f_rest"], params["opacity"], params["scaling"], params["rotation"])
    want = po.activate(*(raw[k].numpy() for k in ("f_dc", "f_rest", "opacity", "scaling", "rotation")))
    for k, a, w in zip(ACT, acts, want):
        assert rel(a.detach().cpu().numpy(), w) < TOL, k
    up = {k: v.cuda() for k, v in mpg.make_upstream(c["P"], c["M"], c["seed"]).items()}
    # only the SH features and the rotations receive a gradient
    torch.autograd.backward([acts[0], acts[3]], [up["shs"], up["rotations"]])
    gw = po.activate_grad(raw["opacity"].numpy(), raw["scaling"].numpy(), raw["rotation"].numpy(), up["shs"].cpu().numpy(),
                          np.zeros((c["P"], 1)), np.zeros((c["P"], 3)), up["rotations"].cpu().numpy())
    assert rel(params["f_dc"].grad.cpu().numpy(), gw[0]) < TOL
    assert params["f_rest"].grad is None
    assert rel(params["rotation"].grad.cpu().numpy(), gw[4]) < TOL
    for k in ("opacity", "scaling"):
        assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0


@pytest.mark.gpu
def test_fused_adam_skips_parameters_without_gradient_and_handles_odd_sizes():
    """A group whose parameter has no .grad keeps its state untouched (torch semantics);
    sizes that are not multiples of 4 and unaligned views take the scalar path."""
    from binocular3dgs_b200 import parameters
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1001 * 3 + 1, generator=g).cuda()
    pa = [torch.nn.Parameter(base[1:].clone().view(1001, 3)), torch.nn.Parameter(torch.randn(7, generator=g).cuda()),
          torch.nn.Parameter(torch.randn(5, 4, generator=g).cuda())]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = torch.optim.Adam([{"params": [p], "lr": 0.01 * (i + 1)} for i, p in enumerate(pa)], lr=0.0, eps=1e-15)
    ob = parameters.FusedAdam([{"params": [p], "lr": 0.01 * (i + 1)} for i, p in enumerate(pb)], lr=0.0, eps=1e-15)
    for step in range(3):
        for plist in (pa, pb):
            gen = torch.Generator().manual_seed(50 + step)
            for i, p in enumerate(plist):
                p.grad = None if (i == 1 and step != 1) else torch.randn(p.shape, generator=gen).cuda()
        oa.step()
        ob.step()
    for a, b in zip(pa, pb):
        assert float((a - b).abs().max()) < 1e-6
    assert float(oa.state[pa[1]]["step"]) == float(ob.state[pb[1]]["step"]) == 1.0
