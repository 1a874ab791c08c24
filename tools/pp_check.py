"""Check + time the attention forward (run once with XCLIP_ATTN_PP=1 and once without)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from x_clip_b200 import kernels as K
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_attention import _ref_attention, _mk

dev = torch.device("cuda:0")
tag = ("tail" if os.environ.get("XCLIP_ATTN_TAIL") == "1" else "notail") + "/v" + os.environ.get("XCLIP_ATTN_PP_VARIANT", "0") + ("/bwd16" if os.environ.get("XCLIP_ATTN_BWD16") == "1" else "")
ok = True
for (B, n, H, masked) in [(2, 129, 2, True), (40, 129, 4, False), (2, 145, 3, False), (2, 197, 12, False), (2, 257, 8, True),
                          (1, 320, 2, True), (40, 257, 8, True), (30, 197, 12, False), (50, 300, 4, True),
                          (64, 160, 5, True)]:
    qkv, mask = _mk(B, n, H, masked, dev)
    o, lse = K.attn_fwd(qkv, mask, B, n, H, 0.125)
    torch.cuda.synchronize()
    ref, s = _ref_attention(qkv, mask, B, n, H, 0.125)
    err = (o.float() - ref).abs().max().item()
    lse_ref = torch.logsumexp(s, -1) * 1.4426950408889634
    lerr = (lse - lse_ref).abs().max().item()
    good = err < 2e-2 and lerr < 2e-2
    ok &= good
    print(f"[{tag}] B={B} n={n} H={H} masked={masked} max|do|={err:.4f} max|dlse|={lerr:.4f} {'OK' if good else 'FAIL'}")
for (B, n, H) in [(1024, 257, 8), (1024, 197, 12), (512, 320, 8)]:
    qkv, mask = _mk(B, n, H, False, dev)
    for _ in range(3): K.attn_fwd(qkv, mask, B, n, H, 0.125)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): K.attn_fwd(qkv, mask, B, n, H, 0.125)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * B * H * n * n * 64
    print(f"[{tag}] time B={B} n={n} H={H}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s")
for (B, n, H) in ([(1024, 257, 8)] if os.environ.get("XCLIP_ATTN_PP_VARIANT", "0") == "0" else []):
    qkv, mask = _mk(B, n, H, False, dev)
    o, lse = K.attn_fwd(qkv, mask, B, n, H, 0.125)
    d_o = torch.randn(B * n, H * 64, device=dev).bfloat16()
    for _ in range(3): K.attn_bwd(qkv, mask, o, d_o, lse, B, n, H, 0.125)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): K.attn_bwd(qkv, mask, o, d_o, lse, B, n, H, 0.125)
    e1.record(); torch.cuda.synchronize()
    print(f"[{tag}] bwd time B={B} n={n} H={H}: {e0.elapsed_time(e1) / 10:.3f} ms")
print("ALL OK" if ok else "SOME FAILED")
