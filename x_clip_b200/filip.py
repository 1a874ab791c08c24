"""FILIP fine-grained loss (reference x_clip/x_clip.py:799-811 + :821-847) on the CUDA kernels.

Schedule (see csrc/filip.cu): two segment-max GEMM passes (text tokens x image tokens and the
transposed orientation, fp32-accurate through the split-bf16 operands), two tiny reductions to the
[B,B] similarity matrices, row-wise InfoNCE/DCL; backward re-expands the saved argmax into a
one-hot weighted bf16 operand chunk by chunk and runs two plain tcgen05 GEMMs per chunk.
The 6-D similarity tensor of the reference is never materialised.  Single process only: the
reference cannot all-gather FILIP latents either."""
from __future__ import annotations

import torch

from . import kernels as K
from . import distributed as D_
from .clip import Unsupported

F32 = torch.float32


class FilipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zt, zi, zt_x, zi_x, temperature, ops, text_mask, dcl: bool):
        extra = len(ops) == 4
        B, T, D = zt.shape
        I = zi.shape[1]
        dev = zt.device
        temp_exp = temperature.detach().float().exp().reshape(1)
        mask = text_mask.reshape(B * T)
        cnt = text_mask.sum(dim=1).clamp(min=1e-6).float()                       # x_clip.py:40-44
        w_t = (text_mask.float() / cnt[:, None]).reshape(B * T).contiguous()
        w_i = torch.full((B * I,), 1.0 / I, device=dev, dtype=F32)
        col_mul = mask.float().contiguous()
        col_add = torch.where(mask, torch.zeros((), device=dev),
                              torch.full((), -torch.finfo(F32).max, device=dev)).contiguous()

        (trow, tcol), (irow, icol) = ops[0], ops[1]
        m1, a1 = K.filip_segmax(trow, icol, temp_exp, I, None, None)             # [B*T, B]
        t2i = K.filip_reduce(m1, w_t, B, T, B, transpose=False)
        if extra:
            (txrow, txcol), (ixrow, ixcol) = ops[2], ops[3]
            m2, a2 = K.filip_segmax(ixrow, txcol, temp_exp, T, col_mul, col_add)  # [B*I, B]
        else:
            m2, a2 = K.filip_segmax(irow, tcol, temp_exp, T, col_mul, col_add)
        i2t = K.filip_reduce(m2, w_i, B, I, B, transpose=True)                   # [text, image]

        loss = torch.zeros(1, device=dev, dtype=F32)
        lse_t = K.filip_nce_fwd(t2i, dcl, loss, 1.0 / (2 * B))
        lse_i = K.filip_nce_fwd(i2t, dcl, loss, 1.0 / (2 * B))

        ctx.cfg = (dcl, extra, B, T, I, D)
        ctx.stuff = (temp_exp, ops, w_t, w_i, m1, a1, m2, a2, t2i, i2t, lse_t, lse_i)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gloss):
        dcl, extra, B, T, I, D = ctx.cfg
        temp_exp, ops, w_t, w_i, m1, a1, m2, a2, t2i, i2t, lse_t, lse_i = ctx.stuff
        dev = gloss.device
        gscale = (gloss.detach().float() / (2.0 * B)).reshape(1)
        g_t = K.filip_nce_bwd(t2i, lse_t, dcl, gscale)                           # [x, y]
        g_i = K.filip_nce_bwd(i2t, lse_i, dcl, gscale).t().contiguous()          # -> [y, x]
        dtemp = torch.zeros(1, device=dev, dtype=F32)

        def one_pass(seg_arg, seg_max, wmat, rowscale, rows_per_sample, seg_len, row_ops, col_ops):
            """rows x cols pass: returns (d rows [R, D], d cols [C, D]) in fp32."""
            R = seg_arg.shape[0]
            C = B * seg_len
            row_hi = row_ops[0][:, :D]
            col_hi = col_ops[1][:, :D]
            d_rows = torch.empty((R, D), device=dev, dtype=F32)
            d_cols = torch.zeros((C, D), device=dev, dtype=F32)
            chunk = max(128, min(R, (1 << 29) // max(C, 1)) // 128 * 128)
            for r0 in range(0, R, chunk):
                rows = min(chunk, R - r0)
                g = K.filip_expand(seg_arg, seg_max, wmat, rowscale, temp_exp, r0, rows,
                                   rows_per_sample, seg_len, B, dtemp)
                K.gemm(g, col_hi, b_major=1, out=d_rows[r0:r0 + rows])
                K.gemm(g, row_hi[r0:r0 + rows], a_major=1, b_major=1, out=d_cols, accumulate=True)
            return d_rows, d_cols

        t_ops, i_ops = ops[0], ops[1]
        dzt, dzi = one_pass(a1, m1, g_t, w_t, T, I, t_ops, i_ops)                # rows = text tokens
        if extra:
            dzi_x, dzt_x = one_pass(a2, m2, g_i, w_i, I, T, ops[3], ops[2])      # rows = image tokens
        else:
            dzi_b, dzt_b = one_pass(a2, m2, g_i, w_i, I, T, i_ops, t_ops)
            dzt += dzt_b
            dzi += dzi_b
            dzt_x = dzi_x = None
        ctx.stuff = None
        shp_t, shp_i = (B, T, D), (B, I, D)
        return (dzt.view(shp_t), dzi.view(shp_i),
                None if dzt_x is None else dzt_x.view(shp_t),
                None if dzi_x is None else dzi_x.view(shp_i),
                dtemp.reshape(()), None, None, None)


def filip_loss(clip, zt, zi, zt_x, zi_x, ops, text_mask):
    if clip.requires_all_gather and D_.world()[1] > 1:
        raise Unsupported("x_clip_b200: use_all_token_embeds with world_size > 1 is not implemented "
                          "(the reference cannot gather FILIP latents either)")
    T, I = zt.shape[1], zi.shape[1]
    if T % 16 or I % 16 or T > 256 or I > 256:
        raise Unsupported(f"x_clip_b200: FILIP needs token counts that are multiples of 16 and <= 256 "
                          f"(got {T} text, {I} image tokens)")
    extra = clip.extra_latent_projection
    return FilipLossFn.apply(zt, zi, zt_x if extra else None, zi_x if extra else None,
                             clip.temperature, tuple(ops), text_mask.contiguous(),
                             clip.decoupled_contrastive_learning)
