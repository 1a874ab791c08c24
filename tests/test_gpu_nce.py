"""Logits + InfoNCE/DCL kernels vs an fp32 torch restatement of x_clip/x_clip.py:813-847."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _unit(n, d, dev, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
    return z.to(dev).bfloat16()


@pytest.mark.parametrize("R,C,off,D,dcl,temp_exp", [
    (6, 6, 0, 256, False, 2.5), (128, 128, 0, 512, True, 2.5), (200, 1000, 300, 512, False, 2.5),
    (1024, 1024, 0, 512, True, 2.5), (384, 3000, 1500, 512, False, 2.5),
    # CLIP-style logit scale: exp(temperature) = 100.  A fixed shift exp(s - 100) underflows every
    # row whose cosines stay below ~0 (random latents!); the online row maximum must not.
    (300, 2000, 700, 512, False, 100.0), (128, 128, 0, 512, True, 100.0)])
def test_nce_fwd_bwd(cuda_device, R, C, off, D, dcl, temp_exp):
    from x_clip_b200 import kernels as K
    dev = cuda_device
    b = _unit(C, D, dev, 1)
    a = _unit(C, D, dev, 2)[off:off + R].contiguous()
    temp = torch.tensor([temp_exp], device=dev)
    lse, pos = K.nce_fwd(a, b, temp, off, dcl)
    s = (temp.double() * (a.double() @ b.double().t()))
    eye = torch.zeros(R, C, dtype=torch.bool, device=dev)
    eye[torch.arange(R), torch.arange(R) + off] = True
    ref_lse = torch.logsumexp(s.masked_fill(eye, -float("inf")) if dcl else s, dim=-1).float()
    s = s.float()
    assert torch.isfinite(lse).all()
    assert torch.allclose(lse, ref_lse, atol=2e-4 * max(1.0, temp_exp / 2.5), rtol=1e-5), (lse - ref_lse).abs().max()
    assert torch.allclose(pos, s[eye], atol=1e-4 * max(1.0, temp_exp / 2.5))

    # backward: g = gs*(w_row*exp(s-lse_row) + w_col*exp(s-lse_col) - w_diag*diag), out = temp*g
    lse_col = torch.randn(C, device=dev) * 0.1 + ref_lse.mean()
    gs = torch.tensor([0.37], device=dev)
    dtemp = torch.zeros(1, device=dev)
    g = K.nce_bwd(a, b, temp, off, dcl, lse, lse_col, 1.0, 0.5, 2.0, gs, dtemp)
    keep = ~eye if dcl else torch.ones_like(eye)
    ref_g = gs * ((torch.exp(s - ref_lse[:, None]) + 0.5 * torch.exp(s - lse_col[None, :])) * keep
                  - 2.0 * eye)
    got = g[:, :C].float()
    err = (got - temp * ref_g).abs().max().item()
    assert err <= 1e-2 * (temp * ref_g).abs().max().item(), err
    assert (g[:, C:] == 0).all()
    ref_dtemp = (ref_g * s).sum().item()
    assert abs(dtemp.item() - ref_dtemp) <= 2e-3 * max(1.0, abs(ref_dtemp)) + 1e-3 * ref_g.abs().sum().item() * 0.01
