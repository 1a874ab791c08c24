#!/usr/bin/env python
"""Benchmark of the CLIP contrastive training hot path (fwd + bwd), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): image-text pairs/s, fwd+bwd.  A "step" is one
`loss = clip(text, images, return_loss=True); loss.backward()` on one batch of synthetic
data (text = randint, images = randn, reference default init).  Workload at any N: cfg2 of
BASELINE.json - the README model (dim 512, 6+6 layers, 8 heads, seq 256, 256px / patch 32,
reference-default visual_patch_dropout 0.5), 1024 pairs per GPU, plain InfoNCE, weak scaling
(global batch 1024*N, negatives all-gathered across ranks).

One JSON line on rank 0: `value` = pairs/s with inputs already on the device; `e2e` = the same
step through the public CLIP.forward with the batch copied from pinned host memory inside the
timed region (double-buffered on a side stream) and the loss read back to the host each step.
`roofline` describes the dominant kernel family (the tcgen05 GEMM), measured live with CUDA
events around every C-ABI call in an extra instrumented step.  `cpu_baseline` / `--impl
reference` time the oracle (fp32 CPU restatement of the reference) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

README_CFG = dict(dim_text=512, dim_image=512, dim_latent=512, num_text_tokens=10000,
                  text_enc_depth=6, text_seq_len=256, text_heads=8, visual_enc_depth=6,
                  visual_image_size=256, visual_patch_size=32, visual_heads=8)
# cfg3 of BASELINE.json: ViT-B/16 image tower + 12-layer text tower (CLIP tokenizer vocabulary)
VITB16_CFG = dict(dim_text=512, dim_image=768, dim_latent=512, num_text_tokens=49408,
                  text_enc_depth=12, text_seq_len=77, text_heads=8, visual_enc_depth=12,
                  visual_image_size=224, visual_patch_size=16, visual_heads=12)
METRIC = "image-text pairs/sec (fwd+bwd)"
WORKLOADS = {
    "cfg2": (README_CFG, "cfg2: README CLIP (dim 512, text 6L seq 256, ViT 6L 256px/32, 8 heads, "
                         "visual_patch_dropout {pd} (reference default 0.5), plain InfoNCE), {b} pairs/GPU"),
    "cfg3": (VITB16_CFG, "cfg3: ViT-B/16 (dim 768, 12L, 12 heads, 224px/16) + text 12L dim 512 seq 77, "
                         "dim_latent 512, visual_patch_dropout {pd}, plain InfoNCE, {b} pairs/GPU"),
}
WORKLOAD = WORKLOADS["cfg2"][1]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="pairs per GPU")
    ap.add_argument("--patch-dropout", type=float, default=0.5)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--loss", default="nce", choices=["nce", "dcl_extra", "filip"],
                    help="nce: plain InfoNCE (BASELINE cfg2/3); dcl_extra: decoupled loss + extra latent "
                         "projections (cfg5); filip: use_all_token_embeds (cfg4)")
    ap.add_argument("--microbatch", type=int, default=0,
                    help="encoder micro-batch (GradCache-style step) - lets --batch 4096 fit one GPU")
    ap.add_argument("--grad-sync", action="store_true",
                    help="N>1: also all-reduce the weight gradients inside the timed step "
                         "(x_clip_b200.distributed.GradSync, buckets overlapped with backward); the "
                         "reference leaves this to the user's DDP wrapper, so it is off by default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the host->device end-to-end leg (large per-GPU batches: no pinned staging)")
    return ap.parse_args()


# --------------------------------------------------------------------------- CPU arm

def cpu_reference_rate(batch: int, steps: int, warmup: int, budget_s: float):
    """pairs/s of the oracle (fp32 CPU restatement of the reference's CLIP.forward + backward)
    on all host cores, README model, `batch` pairs per step, default patch dropout 0.5."""
    import torch
    from oracle import clip_oracle as O
    # all host threads up to 32: beyond that torch's intra-op parallelism degrades on these small
    # per-step shapes (measured on the 128-thread GPU host: 49 s/step with 128 threads)
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    cfg = O.ClipConfig(**README_CFG)
    state = O.protocol_state_dict(cfg, 1234)
    p = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    n_patch = (cfg.visual_image_size // cfg.visual_patch_size) ** 2
    times = []
    t_start = time.time()
    for i in range(warmup + steps):
        text, image = O.protocol_inputs(cfg, batch, 100 + i)
        keep = torch.randn(batch, n_patch).topk(n_patch // 2, dim=-1).indices
        t0 = time.perf_counter()
        loss = O.clip_forward(p, text, image, cfg, keep=keep)
        loss.backward()
        dt = time.perf_counter() - t0
        for v in p.values():
            v.grad = None
        if i >= warmup:
            times.append(dt)
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    ms = 1e3 * sum(times) / len(times)
    return batch / (ms / 1e3), ms, cores, len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 8
    rate, ms, cores, n = cpu_reference_rate(batch, args.steps, args.warmup, budget_s=150.0)
    sample = f"{n} steps of {batch} pairs (fwd+bwd, fp32, README model, patch dropout 0.5)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD.format(b=args.batch, pd=0.5), "sample_batch": batch,
                   "note": "CPU arm: oracle port of the reference (pure-PyTorch fp32), all host threads"},
        "cpu_baseline": {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": rate, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks

class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        busy = [x for x in sm if x > 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None,
                "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- GPU arm

def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - x_clip_b200 has no CPU path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    t_start = time.time()

    def note(msg):
        if os.environ.get("XCLIP_BENCH_VERBOSE"):
            print(f"[bench rank {rank} +{time.time() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    import x_clip_b200
    from x_clip_b200 import _lib, kernels
    lib = _lib.load()
    _lib.call("xclip_init")

    B = args.batch
    torch.manual_seed(0)
    model_cfg, workload_txt = WORKLOADS[args.workload]
    loss_kw = {"nce": {}, "dcl_extra": dict(decoupled_contrastive_learning=True, extra_latent_projection=True),
               "filip": dict(use_all_token_embeds=True)}[args.loss]
    clip = x_clip_b200.CLIP(**model_cfg, **loss_kw, visual_patch_dropout=args.patch_dropout,
                            microbatch=args.microbatch or None).to(dev)
    clip.train()
    params = [p for p in clip.parameters()]

    # synthetic host batches in pinned memory (two alternating buffers) + device-resident copy
    g = torch.Generator().manual_seed(1 + rank)
    host = []
    for _ in range(0 if args.no_e2e else 2):
        t = torch.randint(0, model_cfg["num_text_tokens"], (B, model_cfg["text_seq_len"]),
                          generator=g).pin_memory()
        im = torch.randn(B, 3, model_cfg["visual_image_size"], model_cfg["visual_image_size"],
                         generator=g).pin_memory()
        host.append((t, im))
    if args.no_e2e:
        gd = torch.Generator(device=dev).manual_seed(1 + rank)
        dev_text = torch.randint(0, model_cfg["num_text_tokens"], (B, model_cfg["text_seq_len"]),
                                 generator=gd, device=dev)
        dev_img = torch.randn(B, 3, model_cfg["visual_image_size"], model_cfg["visual_image_size"],
                              generator=gd, device=dev)
        h2d_bytes = 0
    else:
        dev_text = host[0][0].to(dev)
        dev_img = host[0][1].to(dev)
        h2d_bytes = host[0][0].numel() * 8 + host[0][1].numel() * 4

    grad_sync = None
    if args.grad_sync and world > 1:
        from x_clip_b200.distributed import GradSync
        grad_sync = GradSync(clip)

    def step(text, image):
        for p in params:
            p.grad = None
        loss = clip(text, image, return_loss=True)
        loss.backward()
        if grad_sync is not None:
            grad_sync.finish()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    note("model + data ready")
    # ---- warm-up
    for i in range(max(args.warmup, 1)):
        step(dev_text, dev_img)
        note(f"warm-up step {i} enqueued")
    barrier()
    note("warm-up done")

    # ---- (1) device-resident timing
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.xclip_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(dev_text, dev_img)
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    note(f"device-resident timing done: {ms_dev / args.steps:.1f} ms/step")
    launches = lib.xclip_launch_count()
    last_loss = loss.item()
    clocks = sampler.stop() if rank == 0 else None

    # ---- (2) end-to-end timing: pinned host -> device copy of every step's batch (side stream,
    #          double-buffered) + loss read-back, all inside the timed region
    ms_e2e = None
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [] if args.no_e2e else [(torch.empty_like(dev_text), torch.empty_like(dev_img)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            slots[s][0].copy_(host[s][0], non_blocking=True)
            slots[s][1].copy_(host[s][1], non_blocking=True)
            ready[s].record(copy_stream)

    if not args.no_e2e:
        for s in range(2):
            consumed[s].record(torch.cuda.current_stream())
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        issue_copy(0)
        for i in range(args.steps):
            s = i % 2
            if i + 1 < args.steps:
                issue_copy(i + 1)
            torch.cuda.current_stream().wait_event(ready[s])
            loss = step(slots[s][0], slots[s][1])
            consumed[s].record(torch.cuda.current_stream())
            _ = loss.item()                      # device -> host read of the step's result
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))

    note("e2e timing done")
    # ---- (3) instrumented step for the roofline (rank 0 only, after the timed regions)
    prof = None
    if not args.no_profile:        # every rank runs the step (it contains collectives); rank 0 records
        if rank == 0:
            kernels.PROF.start()
        step(dev_text, dev_img)
        if rank == 0:
            prof = kernels.PROF.stop()
    barrier()

    note("profile step done")
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PF sustained (B200_PROFILING.md)"

    roofline = None
    families = None
    if prof:
        total_ms = sum(d["ms"] for d in prof.values())
        families = {k: {"calls": d["calls"], "ms": round(d["ms"], 3),
                        "share": round(d["ms"] / total_ms, 4),
                        "tflops": round(d["flops"] / d["ms"] / 1e9, 1) if d["flops"] else None,
                        "gbs": round(d["bytes"] / d["ms"] / 1e6, 1)} for k, d in sorted(prof.items())}
        gem = [prof[k] for k in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad") if k in prof]
        g_ms = sum(d["ms"] for d in gem)
        g_fl = sum(d["flops"] for d in gem)
        g_calls = sum(d["calls"] for d in gem)
        achieved = g_fl / g_ms / 1e9      # TFLOP/s
        # DRAM traffic per launch: algorithmic bytes of the average launch (counted live) times
        # the traffic/algorithmic ratio measured by `ncu --set full` on representative launches
        # (profiles/r1_ncu_summary.md); null when no capture is committed
        traffic = None
        tf = ROOT / "profiles" / "gemm_traffic.json"
        if tf.exists():
            ratio = json.loads(tf.read_text()).get("traffic_over_algorithmic")
            if ratio:
                traffic = round(ratio * sum(d["bytes"] for d in gem) / g_calls)
        roofline = {"kernel": "gemm_bf16_kernel (tcgen05, fwd+dgrad+wgrad launches)",
                    "bound": "tensor", "achieved": round(achieved, 1), "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": round(achieved / peak_tf, 4), "traffic": traffic,
                    "peak_source": peak_src, "launches_per_step": g_calls,
                    "avg_launch_ms": round(g_ms / g_calls, 4),
                    "flops_per_launch": g_fl / g_calls,
                    "algorithmic_bytes_per_launch": round(sum(d["bytes"] for d in gem) / g_calls),
                    "share_of_step": round(g_ms / total_ms, 4)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "cfg2":
        rate, ms, cores, n = cpu_reference_rate(4, steps=30, warmup=2, budget_s=20.0)
        cpu_baseline = {"value": round(rate, 3), "unit": "pairs/s", "cores": cores, "kind": "port",
                        "sample": f"{n} steps of cfg1 (README model, batch 4, fp32, fwd+bwd) on the "
                                  f"oracle port of the reference, {ms:.0f} ms/step"}

    Bg = B * world
    line = {
        "metric": METRIC, "value": round(Bg * args.steps / (ms_dev / 1e3), 2), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_txt.format(b=B, pd=args.patch_dropout).replace(
                       "plain InfoNCE", {"nce": "plain InfoNCE", "dcl_extra": "DCL + extra latent projection",
                                         "filip": "FILIP (use_all_token_embeds)"}[args.loss]) + (f", encoder micro-batch {args.microbatch} "
                   "(two-pass GradCache step: +1 encoder forward)" if args.microbatch else ""),
                   "global_batch": Bg, "parallelism": f"dp{world}" + ("+grad-allreduce" if args.grad_sync and world > 1 else ""),
                   "l2": "per-step working set (tens of GB of activations) >> 126 MB L2; no flush needed",
                   "timing": "CUDA events on the launching stream, barrier+synchronize both sides, max over ranks",
                   "loss": round(last_loss, 5)},
        "clocks": clocks,
        "e2e": None if ms_e2e is None else {
            "value": round(Bg * args.steps / (ms_e2e / 1e3), 2), "unit": "pairs/s",
            "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
            "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "kernel_families": families,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
