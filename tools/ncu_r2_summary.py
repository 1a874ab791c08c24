"""gpurun_out/prof_<name>_{details,raw}.csv (tools/r2_profile.sh) -> profiles/r2_ncu_summary.md

    python tools/ncu_r2_summary.py > profiles/r2_ncu_summary.md"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = PEAKS.get("hbm_gbs", 6500.0)
TF = PEAKS.get("bf16_tflops", 1640.0)
B, n, H = 512, 98, 12
M, d = 50176, 768
R, C, D = 4096, 32768, 1536
# name -> (description, algorithmic flops, algorithmic bytes)
ALG = {
    "attn_small_fwd": (f"attention fwd B={B} n={n} h={H} (cfg3 image tower micro-batch)", 4.0 * B * H * n * n * 64, 2.0 * B * n * H * 64 * 4),
    "attn_small_bwd": (f"attention bwd B={B} n={n} h={H}", 10.0 * B * H * n * n * 64, 2.0 * B * n * H * 64 * 8),
    "attn_wg_fwd": ("attention fwd B=1024 n=257 h=8 (cfg2 text tower)", 4.0 * 1024 * 8 * 257 * 257 * 64, 2.0 * 1024 * 257 * 512 * 4),
    "attn_big_bwd": ("attention bwd B=1024 n=257 h=8", 10.0 * 1024 * 8 * 257 * 257 * 64, 2.0 * 1024 * 257 * 512 * 8),
    "nce_fwd": (f"logits + InfoNCE fwd, {R} local rows x {C} columns, D = 3*512 (global batch 32768)", 2.0 * R * C * D, 2.0 * (R + C) * D + 8 * R),
    "nce_bwd": ("logits + InfoNCE bwd (bf16 g[4096, 32768] by TMA stores)", 2.0 * R * C * D, 2.0 * (R + C) * D + 2.0 * R * C),
    "pair_store": (f"CTA-pair GEMM [{M}x{d}]x[{8 * d}x{d}]^T (FF-up shape, plain epilogue)", 2.0 * M * 8 * d * d, 2.0 * (M * d + 8 * d * d + M * 8 * d)),
    "ff_up": ("fused FF-up: GEMM + GEGLU epilogue (u 8d + hp 4d written)", 2.0 * M * 8 * d * d, 2.0 * (M * d + 8 * d * d + M * 12 * d)),
    "ff_down": ("fused FF-down: GEMM + LayerNorm fold + residual", 2.0 * M * 4 * d * d, 2.0 * (M * 4 * d + 4 * d * d + 3 * M * d)),
    "ff_bwd": ("fused FF backward: dgrad GEMM + LayerNorm/GEGLU backward epilogue (TMA-pipelined u, variant 1)", 2.0 * M * 4 * d * d, 2.0 * (M * d + 4 * d * d + 16 * M * d)),
    "pair_wgrad": (f"CTA-pair wgrad [{8 * d}x{M}]x[{M}x{d}] (FF-up weight gradient, split-K, fp32 red.add)", 2.0 * M * 8 * d * d, 2.0 * (M * 8 * d + M * d) + 4.0 * 8 * d * d),
    "pair_dgrad": (f"CTA-pair dgrad [{M}x{d}]x[{d}x{4 * d}] (MN-major B operand; the un-fused FF-down input gradient)", 2.0 * M * 4 * d * d, 2.0 * (M * d + 4 * d * d + M * 4 * d)),
}


def details(name):
    p = OUT / f"prof_{name}_details.csv"
    if not p.exists():
        return None
    rows = list(csv.reader(open(p)))
    hdr = rows[0]
    mi, vi, ui = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    out = {}
    for r in rows[1:]:
        out.setdefault(r[mi], (r[vi], r[ui]))
    return out


def raw(name):
    p = OUT / f"prof_{name}_raw.csv"
    rows = list(csv.reader(open(p)))
    return dict(zip(rows[0], rows[-1])), dict(zip(rows[0], rows[1]))


def num(x):
    return float(str(x).replace(",", ""))


def main():
    print("# ncu `--set full` captures, round 2 (B200; `tools/r2_profile.sh`, one launch per kernel)\n")
    print("Captured with `--clock-control none`, one kernel at a time: times are cold-cache / serialised and")
    print(f"clocks are higher than inside a full step. Peaks (MEASURED_PEAKS.json): HBM {HBM} GB/s, bf16 {TF} TFLOP/s burst.\n")
    print("| kernel | launch | time us | TFLOP/s (alg.) | frac of bf16 burst peak | tensor-pipe active % | DRAM bytes GB (ncu) | algorithmic GB | DRAM GB/s | frac of HBM peak | regs | binding roofline |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for name, (desc, flops, nbytes) in ALG.items():
        dt = details(name)
        if dt is None:
            continue
        r, units = raw(name)
        dur_v, dur_u = dt["Duration"]
        us = num(dur_v) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(dur_u, 1.0)
        scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
        dram = sum(num(r[k]) * scale.get(units[k], 1e-9) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in r)
        tens = r.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "")
        kern = next(iter(csv.DictReader(open(OUT / f"prof_{name}_details.csv"))))["Kernel Name"].split("(")[0]
        tfs = flops / us / 1e6
        gbs = dram / us * 1e6
        ft, fh = tfs / TF, gbs / HBM
        print(f"| `{kern}` | {desc} | {us:.1f} | {tfs:.0f} | {ft:.2f} | {num(tens):.1f} | {dram:.3f} | {nbytes / 1e9:.3f} | "
              f"{gbs:.0f} | {fh:.2f} | {dt['Registers Per Thread'][0]} | {'HBM' if fh > ft else 'tensor'} |")


if __name__ == "__main__":
    sys.exit(main())
