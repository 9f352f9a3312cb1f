"""Which float32 evaluation order does torch use for F.normalize / sigmoid / exp on CUDA?
Counts bit mismatches between torch's activations and candidate formulas (each torch op rounds
once, so a chain of torch ops reproduces an explicit rounding sequence), and between torch and
this library's activation kernel (csrc/common.cuh: act_scale, act_opacity, act_rotation — the
same device functions the raw-parameter entry fuses into the preprocess).
Result on torch 2.11 / B200: exp and sigmoid bit-identical; ||q||^2 = (x0^2 + x2^2) + (x1^2 + x3^2)."""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
P = 1_000_000
q = (torch.randn(P, 4, generator=g) * 1.3).to(dev)
nrm_t = torch.linalg.vector_norm(q, dim=1)
ref = torch.nn.functional.normalize(q)
x = [q[:, i].contiguous() for i in range(4)]
sq = [v * v for v in x]


def fma(a, b, c):   # float32 fma through float64 (exact product)
    return (a.double() * b.double() + c.double()).float()


def bad(n2):
    return int((n2.sqrt().view(torch.int32) != nrm_t.view(torch.int32)).sum())


found = []
for p in itertools.permutations(range(4)):
    a, b, c, d = p
    for name, n2 in (("((a+b)+c)+d", ((sq[a] + sq[b]) + sq[c]) + sq[d]), ("(a+b)+(c+d)", (sq[a] + sq[b]) + (sq[c] + sq[d])),
                     ("fma chain", fma(x[d], x[d], fma(x[c], x[c], fma(x[b], x[b], sq[a]))))):
        n = bad(n2)
        if n == 0:
            found.append((name, p))
print("orders of the squares that reproduce torch.linalg.vector_norm bit for bit:", found)
print("F.normalize == q / clamp_min(norm, 1e-12):",
      int(((q / nrm_t.clamp_min(1e-12)[:, None]).view(torch.int32) != ref.view(torch.int32)).any(dim=1).sum()) == 0)

from binocular3dgs_b200 import parameters  # noqa: E402
raw_o = (torch.randn(P, 1, generator=g) * 3).to(dev)
raw_s = (torch.randn(P, 3, generator=g) * 2 - 3).to(dev)
f_dc, f_rest = torch.randn(P, 1, 3, generator=g).to(dev), torch.randn(P, 3, 3, generator=g).to(dev)
shs, op, sc, rot = parameters.activate(f_dc, f_rest, raw_o, raw_s, q)
print("native vs torch  sigmoid mismatches:", int((op.view(torch.int32) != torch.sigmoid(raw_o).view(torch.int32)).sum()))
print("native vs torch  exp     mismatches:", int((sc.view(torch.int32) != torch.exp(raw_s).view(torch.int32)).sum()))
print("native vs torch  normalize mismatching rows:", int((rot.view(torch.int32) != ref.view(torch.int32)).any(dim=1).sum()))
print("native vs torch  cat equal:", bool(torch.equal(shs, torch.cat((f_dc, f_rest), dim=1))))
