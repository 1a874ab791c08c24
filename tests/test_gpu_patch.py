"""Patch-embedding front end (xclip_patchify_gather + GEMM with gathered position rows) vs the
reference's patchify -> Linear -> + pos -> PatchDropout order (x_clip/x_clip.py:356-359, :379-385)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_patchify(img, p):
    b, c, H, W = img.shape
    x = img.view(b, c, H // p, p, W // p, p).permute(0, 2, 4, 3, 5, 1)
    return x.reshape(b, (H // p) * (W // p), p * p * c)


@pytest.mark.parametrize("B,C,S,p,drop", [(3, 3, 64, 16, False), (5, 3, 224, 16, True), (2, 3, 256, 32, True),
                                         (2, 1, 64, 8, True)])
def test_patchify_gather_and_embed(cuda_device, B, C, S, p, drop):
    from x_clip_b200 import kernels as K
    from x_clip_b200 import engine as E
    dev = cuda_device
    g = torch.Generator().manual_seed(B + S)
    img = torch.randn(B, C, S, S, generator=g).to(dev)
    n = (S // p) ** 2
    keep = None
    if drop:
        keep = torch.randn(B, n, generator=g).topk(max(1, n // 2), dim=-1).indices.to(dev)
    got = K.patchify_gather(img, p, keep)
    ref = _ref_patchify(img, p)
    if keep is not None:
        ref = torch.gather(ref, 1, keep[:, :, None].expand(-1, -1, ref.shape[-1]))
    assert torch.equal(got.view(B, -1, ref.shape[-1]), ref.bfloat16())

    # embedding + bias + gathered position rows, forward and parameter gradients
    d = 256
    W = ((torch.rand(d, p * p * C, generator=g) * 2 - 1) / (p * p * C) ** 0.5).to(dev).requires_grad_(True)
    bias = (0.02 * torch.randn(d, generator=g)).to(dev).requires_grad_(True)
    pos = torch.randn(n, d, generator=g).to(dev).requires_grad_(True)
    k = n if keep is None else keep.shape[1]
    index = (torch.arange(n, device=dev).repeat(B) if keep is None else keep.reshape(-1)).to(torch.int32)
    tok = E.PatchEmbedFn.apply(got, index, W, bias, pos)
    dy = torch.randn(B * k, d, generator=g).to(dev)
    tok.backward(dy.bfloat16())
    Wr, br, pr = (t.detach().clone().requires_grad_(True) for t in (W, bias, pos))
    full = _ref_patchify(img, p).bfloat16().float() @ Wr.bfloat16().float().t() + br + pr.bfloat16().float()[None]
    if keep is not None:
        full = torch.gather(full, 1, keep[:, :, None].expand(-1, -1, d))
    full = full.reshape(B * k, d)
    assert (tok.float() - full).abs().max().item() <= 2e-2 * full.abs().max().item()
    full.backward(dy.bfloat16().float())
    for name, a, r in (("dW", W.grad, Wr.grad), ("db", bias.grad, br.grad), ("dpos", pos.grad, pr.grad)):
        rel = (a - r).norm().item() / (r.norm().item() + 1e-9)
        assert rel <= 1e-2, (name, rel)
