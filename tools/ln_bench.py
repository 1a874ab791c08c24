"""LayerNorm forward / backward at the bench's chunk shapes, with the grid-size switches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from x_clip_b200 import _lib, kernels as K  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


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


for (M, d) in [(75264, 768), (59904, 512)]:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, d, generator=g).to(dev).bfloat16()
    r = torch.randn(M, d, generator=g).to(dev).bfloat16()
    dy = torch.randn(M, d, generator=g).to(dev).bfloat16()
    gain = torch.ones(d, device=dev)
    _, st, _, _ = K.layernorm_fwd(x, gain)
    dg = torch.zeros(d, device=dev)
    by = 2.0 * M * d
    for fb in (0, 2, 3, 4, 6):
        lib.xclip_tune_set(3, fb)
        a = timeit(lambda: K.layernorm_fwd(x, gain))
        b = timeit(lambda: K.layernorm_fwd(x, gain, res=r, g2=gain))
        print(f"M={M} d={d} ln_fwd blocks/SM cap {fb or 8}: plain {a*1e3:.1f} us ({2*by/a/1e6:.0f} GB/s) | "
              f"res+chained {b*1e3:.1f} us ({4*by/b/1e6:.0f} GB/s)", flush=True)
    lib.xclip_tune_set(3, 0)
    for bb in (0, 1, 3, 4):
        lib.xclip_tune_set(4, bb)
        a = timeit(lambda: K.layernorm_bwd(dy, x, st, gain, dg=dg))
        b = timeit(lambda: K.layernorm_bwd(dy, x, st, gain, add=r, dg=dg))
        print(f"M={M} d={d} ln_bwd blocks/SM {bb or 2}: plain {a*1e3:.1f} us ({3*by/a/1e6:.0f} GB/s) | "
              f"+add {b*1e3:.1f} us ({4*by/b/1e6:.0f} GB/s)", flush=True)
    lib.xclip_tune_set(4, 0)
