"""CPU check of the algebra behind the attention tail-token path (x_clip_b200/csrc/attention_tail.cu).

For n = 128k+1 the tensor-core backward covers the [0,z) x [0,z) block (z = n-1) and the tail
kernel supplies row z, column z and three scalars per token that the block kernel folds in as
rank-1 updates.  This test restates that decomposition in fp32 torch, exactly as the kernels
compute it (base-2 log-sum-exp, saved lse/delta of the FULL softmax), and compares it with
autograd through the reference attention core (x_clip/x_clip.py:217-244).  It checks the
formulas (signs, scale, the corner element counted once, masked keys) - not the CUDA code.
"""
import math

import pytest
import torch

LOG2E = 1.4426950408889634


def _reference(q, k, v, mask, scale):
    s = (q * scale) @ k.t()
    if mask is not None:
        s = s.masked_fill(~mask[None, :], -torch.finfo(torch.float32).max)
    p = s.softmax(-1)
    return p @ v, s


@pytest.mark.parametrize("n,masked,mask_tail", [(129, False, False), (129, True, False),
                                                (257, True, True), (33, True, True)])
def test_tail_decomposition_matches_autograd(n, masked, mask_tail):
    g = torch.Generator().manual_seed(n + masked)
    d = 64
    scale = d ** -0.5
    q, k, v, d_o = (torch.randn(n, d, generator=g, dtype=torch.float64) for _ in range(4))
    mask = None
    if masked:
        mask = torch.rand(n, generator=g) > 0.3
        mask[0] = True
        mask[n - 1] = not mask_tail
    qa, ka, va = (t.clone().requires_grad_(True) for t in (q, k, v))
    o_ref, s_ref = _reference(qa, ka, va, mask, scale)
    o_ref.backward(d_o)

    z = n - 1
    c = scale * LOG2E
    keep = mask if mask is not None else torch.ones(n, dtype=torch.bool)
    # ---- forward: block rows from the full softmax, row z by the tail formula
    t_all = torch.where(keep[None, :], (q @ k.t()) * c, torch.full((), -torch.finfo(torch.float32).max, dtype=torch.float64))
    m = t_all.max(-1).values
    lse = m + torch.log2(torch.exp2(t_all - m[:, None]).sum(-1))          # base-2, every row
    p_z = torch.exp2(t_all[z] - m[z])
    o_z = (p_z[:, None] * v).sum(0) / p_z.sum()
    assert torch.allclose(o_z, o_ref[z].detach(), atol=1e-10)
    o = o_ref.detach()
    delta = (d_o * o).sum(-1)

    # ---- tensor-core block [0,z) x [0,z): masked keys contribute p = 0 (mul = 0, add = -inf)
    def probs(rows, cols):
        s2 = (q[rows] @ k[cols].t()) * c
        p = torch.exp2(s2 - lse[rows][:, None])
        return torch.where(keep[cols][None, :], p, torch.zeros((), dtype=torch.float64))
    main = torch.arange(z)
    P = probs(main, main)
    dP = d_o[main] @ v[main].t()
    dS = P * (dP - delta[main][:, None]) * scale
    dq = torch.zeros(n, d, dtype=torch.float64)
    dk = torch.zeros_like(dq)
    dv = torch.zeros_like(dq)
    dq[:z] = dS @ k[main]
    dk[:z] = dS.t() @ q[main]
    dv[:z] = P.t() @ d_o[main]

    # ---- tail kernel: column of key z (all queries), row of query z (keys < z)
    allr = torch.arange(n)
    p_c = probs(allr, torch.tensor([z]))[:, 0]
    ds_c = p_c * (d_o @ v[z] - delta) * scale
    p_r = probs(torch.tensor([z]), main)[0]
    ds_r = p_r * (v[main] @ d_o[z] - delta[z]) * scale
    dv[z] = (p_c[:, None] * d_o).sum(0)
    dk[z] = (ds_c[:, None] * q).sum(0)
    dq[z] = (ds_r[:, None] * k[main]).sum(0) + ds_c[z] * k[z]
    # ---- rank-1 terms the block kernel's epilogues add from the three scalars per token
    dq[:z] += ds_c[:z, None] * k[z][None, :]
    dk[:z] += ds_r[:, None] * q[z][None, :]
    dv[:z] += p_r[:, None] * d_o[z][None, :]

    for name, got, ref in (("dq", dq, qa.grad), ("dk", dk, ka.grad), ("dv", dv, va.grad)):
        err = (got - ref).abs().max().item()
        assert err < 1e-9 * max(1.0, ref.abs().max().item()), f"{name}: {err}"
    assert math.isfinite(lse[z].item())
