"""Lock-step numpy emulation of the two tail-token kernels (x_clip_b200/csrc/attention_tail.cu).

Every per-thread variable of the CUDA source becomes an array over the 128 threads of a CTA and
every index expression (base pointers, `inner`, leading dimensions, the dq-workspace slot, the
warp-shuffle reductions) is restated verbatim, operating on the real memory layout
(qkv [B*n, 3*H*64] = q | k | v head-major, d_o / o [B*n, H*64], lse / delta [B, H, n], mask [B, n]).
The block kernel's tail-mode epilogue terms are applied the way attention_bwd.cu applies them
(from the three scalars in the workspace slot).  The result must match autograd through the
reference attention core (x_clip/x_clip.py:217-244).  This pins the index arithmetic of the
CUDA-core path on CPU; it is not a substitute for running the kernels.
"""
import numpy as np
import pytest
import torch

LOG2E = 1.4426950408889634
NT = 128                                  # kTailThreads
FLT_MAX = float(np.finfo(np.float32).max)


def _sum8lanes(x):                        # tail_sum8lanes: butterfly over the 8 lanes of a token group
    return np.repeat(x.reshape(NT // 8, 8).sum(1), 8)


def _block_sum8(a):                       # tail_block_sum8: [NT, 8] per-lane vectors -> 64 values
    # xor 8 / 16 inside a warp: lanes with equal (lane & 7); then the 4 warps through smem
    per_warp = a.reshape(4, 4, 8, 8).sum(1)          # [warp, sub, i]
    tot = per_warp.sum(0)                            # [sub (= lane of warp 0), i]
    return tot.reshape(64)                           # element sub*8 + i  == column lane*8 + i


def emulate_fwd_tail(qkv, ld, mask, o, ldo, lse, B, H, n, scale_log2):
    tid = np.arange(NT)
    grp, sub = tid >> 3, tid & 7
    inner = H * 64
    for bh in range(B * H):
        b, h = bh // H, bh % H
        z = n - 1
        base = b * n * ld + h * 64 + sub * 8                     # element offsets into qkv
        idx8 = np.arange(8)
        qz = qkv[(base + z * ld)[:, None] + idx8]
        m = np.full(NT, -np.inf)
        s_t = np.zeros(320)
        for j0 in range(0, n, 16):
            j = j0 + grp
            valid = j < n
            jj = np.where(valid, j, n - 1)
            kj = qkv[(base + jj * ld + inner)[:, None] + idx8]
            s = _sum8lanes((qz * kj).sum(1))
            keep = np.ones(NT, bool) if mask is None else mask[b * n + jj] != 0
            t = np.where(keep, s * scale_log2, -FLT_MAX)
            w = valid & (sub == 0)
            s_t[j[w]] = t[w]
            m = np.where(valid, np.maximum(m, t), m)
        m = np.full(NT, m.max())                                 # warp_max + cross-warp max
        acc = np.zeros((NT, 8))
        l = np.zeros(NT)
        for it in range((n + 15) // 16):
            j = grp + 16 * it
            act = j < n
            jj = np.where(act, j, 0)
            vj = qkv[(base + jj * ld + 2 * inner)[:, None] + idx8]
            pj = np.where(act, np.exp2(s_t[jj] - m), 0.0)
            l += pj
            acc += pj[:, None] * vj
        L = l.reshape(NT // 8, 8)[:, 0].sum()                    # one lane per group counts once
        out = _block_sum8(acc) / L
        o[(b * n + z) * ldo + h * 64 + np.arange(64)] = out
        lse[(b * H + h) * n + z] = m[0] + np.log2(L)


def emulate_bwd_tail(qkv, ld, mask, d_o, lddo, lse, delta, dqkv, ldg, ws, B, H, n, scale):
    scale_log2 = scale * LOG2E
    tid = np.arange(NT)
    grp, sub = tid >> 3, tid & 7
    inner = H * 64
    idx8 = np.arange(8)
    for bh in range(B * H):
        b, h = bh // H, bh % H
        z = n - 1
        base = b * n * ld + h * 64 + sub * 8
        dbase = b * n * lddo + h * 64 + sub * 8
        lse_bh, delta_bh = (b * H + h) * n, (b * H + h) * n
        qz = qkv[(base + z * ld)[:, None] + idx8]
        kz = qkv[(base + z * ld + inner)[:, None] + idx8]
        vz = qkv[(base + z * ld + 2 * inner)[:, None] + idx8]
        doz = d_o[(dbase + z * lddo)[:, None] + idx8]
        lse_z, delta_z = lse[lse_bh + z], delta[delta_bh + z]
        keep_z = True if mask is None else bool(mask[b * n + z])
        dvz, dkz, dqz = np.zeros((NT, 8)), np.zeros((NT, 8)), np.zeros((NT, 8))
        for t0 in range(0, n, 16):
            valid = t0 + grp < n
            t = np.where(valid, t0 + grp, z)
            qt = qkv[(base + t * ld)[:, None] + idx8]
            kt = qkv[(base + t * ld + inner)[:, None] + idx8]
            vt = qkv[(base + t * ld + 2 * inner)[:, None] + idx8]
            dot = d_o[(dbase + t * lddo)[:, None] + idx8]
            lse_t, delta_t = lse[lse_bh + t], delta[delta_bh + t]
            a = _sum8lanes((qt * kz).sum(1))
            bb = _sum8lanes((dot * vz).sum(1))
            c = _sum8lanes((qz * kt).sum(1))
            d = _sum8lanes((doz * vt).sum(1))
            p_c = np.where(valid & keep_z, np.exp2(a * scale_log2 - lse_t), 0.0)
            ds_c = p_c * (bb - delta_t) * scale
            keep_t = np.ones(NT, bool) if mask is None else mask[b * n + t] != 0
            p_r = np.where((t < z) & keep_t, np.exp2(c * scale_log2 - lse_z), 0.0)
            ds_r = p_r * (d - delta_z) * scale
            dvz += p_c[:, None] * dot
            dkz += ds_c[:, None] * qt
            dqz += ds_r[:, None] * kt
            corner = valid & (t == z)
            dqz += np.where(corner, ds_c, 0.0)[:, None] * kz
            rec = valid & (t != z) & (sub == 0)
            w = (b * n + t[rec]) * inner + h * 64
            ws[w], ws[w + 1], ws[w + 2] = ds_c[rec], ds_r[rec], p_r[rec]
        gz = (b * n + z) * ldg + h * 64
        dqkv[gz + np.arange(64)] = _block_sum8(dqz)
        dqkv[gz + inner + np.arange(64)] = _block_sum8(dkz)
        dqkv[gz + 2 * inner + np.arange(64)] = _block_sum8(dvz)


def _reference(qkv, mask, B, n, H, scale):
    q, k, v = qkv.view(B, n, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q * scale) @ k.transpose(-1, -2)
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :], -torch.finfo(torch.float32).max)
    p = s.softmax(-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * n, H * 64), s


@pytest.mark.parametrize("B,H,n,masked", [(2, 3, 129, True), (1, 2, 129, False), (1, 1, 257, True)])
def test_tail_kernels_emulated_on_the_real_layout(B, H, n, masked):
    g = torch.Generator().manual_seed(n + B)
    scale = 64 ** -0.5
    qkv_t = torch.randn(B * n, 3 * H * 64, generator=g, dtype=torch.float64)
    d_o_t = torch.randn(B * n, H * 64, generator=g, dtype=torch.float64)
    mask_t = None
    if masked:
        mask_t = torch.rand(B, n, generator=g) > 0.3
        mask_t[:, 0] = True
        mask_t[0, n - 1] = False                                  # a masked tail key in batch 0
    leaf = qkv_t.clone().requires_grad_(True)
    o_ref, s_ref = _reference(leaf, mask_t, B, n, H, scale)
    o_ref.backward(d_o_t)
    g_ref = leaf.grad.numpy()
    lse_ref = (torch.logsumexp(s_ref.detach(), -1) * LOG2E).numpy()          # [B,H,n] base 2

    inner, z, nt = H * 64, n - 1, n - 1
    qkv, d_o = qkv_t.numpy().reshape(-1).copy(), d_o_t.numpy().reshape(-1).copy()
    mask = None if mask_t is None else mask_t.numpy().reshape(-1).astype(np.uint8)
    ld, ldo = 3 * inner, inner
    # ---- forward: block rows as the tensor-core kernel leaves them, row z by the tail kernel
    o = o_ref.detach().numpy().reshape(-1).copy()
    lse = lse_ref.reshape(-1).copy()
    for b in range(B):
        o[(b * n + z) * ldo:(b * n + z + 1) * ldo] = np.nan
    lse.reshape(B, H, n)[:, :, z] = np.nan
    emulate_fwd_tail(qkv, ld, mask, o, ldo, lse, B, H, n, scale * LOG2E)
    assert np.allclose(o.reshape(B * n, inner), o_ref.detach().numpy(), atol=1e-9)
    assert np.allclose(lse.reshape(B, H, n), lse_ref, atol=1e-9)

    # ---- backward
    delta = (d_o_t * o_ref.detach()).view(B, n, H, 64).sum(-1).permute(0, 2, 1).contiguous().numpy().reshape(-1)
    dqkv = np.full(B * n * 3 * inner, np.nan)
    ws = np.full(B * n * inner, np.nan)
    emulate_bwd_tail(qkv, ld, mask, d_o, inner, lse, delta, dqkv, 3 * inner, ws, B, H, n, scale)
    # block kernel in tail mode: tiles over [0,nt) x [0,nt), then the rank-1 epilogue terms
    Q = qkv.reshape(B, n, 3, H, 64)
    dO = d_o.reshape(B, n, H, 64)
    G = dqkv.reshape(B, n, 3, H, 64)
    W = ws.reshape(B, n, H, 64)
    L2, D = lse.reshape(B, H, n), delta.reshape(B, H, n)
    for b in range(B):
        keep = np.ones(n, bool) if mask is None else mask.reshape(B, n)[b] != 0
        for h in range(H):
            q, k, v, do = Q[b, :nt, 0, h], Q[b, :nt, 1, h], Q[b, :nt, 2, h], dO[b, :nt, h]
            P = np.exp2((q @ k.T) * scale * LOG2E - L2[b, h, :nt, None]) * keep[None, :nt]
            dS = P * (do @ v.T - D[b, h, :nt, None]) * scale
            kz, qz, doz = Q[b, z, 1, h], Q[b, z, 0, h], dO[b, z, h]
            G[b, :nt, 0, h] = dS @ k + W[b, :nt, h, 0:1] * kz[None, :]       # dQ_t += ds^c_t k_z
            G[b, :nt, 1, h] = dS.T @ q + W[b, :nt, h, 1:2] * qz[None, :]     # dK_t += ds^r_t q_z
            G[b, :nt, 2, h] = P.T @ do + W[b, :nt, h, 2:3] * doz[None, :]    # dV_t += p^r_t dO_z
    got = G.reshape(B * n, 3 * inner)
    assert not np.isnan(got).any(), "some gradient row was never written"
    err = np.abs(got - g_ref).max()
    assert err < 1e-9 * max(1.0, np.abs(g_ref).max()), err
