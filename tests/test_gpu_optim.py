"""Fused AdamW (x_clip_b200.optim, xclip_adamw_step) vs torch.optim.AdamW on the same gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch(cuda_device):
    from x_clip_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(7,), (33, 5), (256, 128), (1,), (1000, 3)]
    ours = [torch.nn.Parameter(torch.randn(s, device=cuda_device)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    opt = FusedAdamW(ours, **kw)
    topt = torch.optim.AdamW(ref, **kw)
    for step in range(5):
        for i, (p, r) in enumerate(zip(ours, ref)):
            if step == 2 and i == 1:            # a parameter without gradient in one step
                p.grad = None
                r.grad = torch.zeros_like(r)
                continue
            g = torch.randn_like(r)
            p.grad = g.clone()
            r.grad = g.clone()
        opt.step()
        topt.step()
    for p, r in zip(ours, ref):
        assert torch.allclose(p, r, atol=2e-6, rtol=1e-5), (p - r).abs().max()
    # the parameters live inside one flat buffer (views), module-visible values are the trained ones
    assert all(p.data_ptr() >= opt.flat.data_ptr() for p in ours)
