"""Does the short-sequence attention kernel suffer from the token-major qkv layout?

A (batch, head) item reads n rows of 128 bytes at a stride of 3*H*64*2 bytes (4.6 KB for ViT-B/16):
every row segment lives in a different DRAM page.  The same kernel run with H = 1 and B*H items sees
the same flops and bytes but each item's q | k | v rows are one contiguous n x 384-byte region.
If the H = 1 run is much faster, a head-major activation layout is worth its plumbing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from x_clip_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (B, n, H) in [(512, 98, 12), (768, 98, 12), (768, 78, 8)]:
    for (b2, h2, tag) in [(B, H, "token-major [B*n, 3*H*64]"), (B * H, 1, "one head per item [B*H*n, 192] (contiguous items)")]:
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(b2 * n, 3 * h2 * 64, generator=g).to(dev).bfloat16()
        d_o = torch.randn(b2 * n, h2 * 64, generator=g).to(dev).bfloat16()
        o, lse = K.attn_fwd(qkv, None, b2, n, h2, 0.125)
        ms_f = timeit(lambda: K.attn_fwd(qkv, None, b2, n, h2, 0.125))
        ms_b = timeit(lambda: K.attn_bwd(qkv, None, o, d_o, lse, b2, n, h2, 0.125))
        by_f = 2.0 * b2 * n * h2 * 64 * 4
        by_b = 2.0 * b2 * n * h2 * 64 * 8
        print(f"B={B} n={n} H={H} {tag}: fwd {ms_f:.4f} ms ({by_f / ms_f / 1e6:.0f} GB/s) | bwd {ms_b:.4f} ms "
              f"({by_b / ms_b / 1e6:.0f} GB/s)", flush=True)
