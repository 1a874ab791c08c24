"""Turn the `ncu --page raw --csv` exports under profiles/ into profiles/r1_ncu_summary.md."""
import csv
import io
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
M, d = 1024 * 257, 512
ALG = {
    ("gemm", 0): ("QKV fwd [263168x512]x[1536x512]^T", 2 * M * d * 1536, 2 * (M * d + 1536 * d + M * 1536)),
    ("gemm", 1): ("FF-up fwd [263168x512]x[4096x512]^T", 2 * M * d * 4096, 2 * (M * d + 4096 * d + M * 4096)),
    ("gemm", 2): ("FF-up dgrad [263168x4096]x[4096x512]", 2 * M * d * 4096, 2 * (M * 4096 + 4096 * d + M * d)),
    ("gemm", 3): ("FF-up wgrad [4096x263168]x[263168x512], split-K + fp32 red", 2 * M * d * 4096, 2 * (M * 4096 + M * d) + 4 * 4096 * d),
    ("attn", 0): ("attention fwd B=1024 n=257 h=8", 4 * 1024 * 8 * 257 * 257 * 64, 2 * 1024 * 257 * 512 * 4),
    ("attn", 1): ("attention bwd B=1024 n=257 h=8", 10 * 1024 * 8 * 257 * 257 * 64, 2 * 1024 * 257 * 512 * 8),
    ("attn", 2): ("attention fwd B=1024 n=33 h=8", 4 * 1024 * 8 * 33 * 33 * 64, 2 * 1024 * 33 * 512 * 4),
    ("attn", 3): ("attention bwd B=1024 n=33 h=8", 10 * 1024 * 8 * 33 * 33 * 64, 2 * 1024 * 33 * 512 * 8),
    ("nce", 0): ("InfoNCE fwd, 4096 local rows x 32768 cols, D=3*512", 2 * 4096 * 32768 * 1536, 2 * (4096 + 32768) * 1536 + 8 * 4096),
    ("nce", 1): ("InfoNCE bwd (writes bf16 g[4096,32768])", 2 * 4096 * 32768 * 1536, 2 * (4096 + 32768) * 1536 + 2 * 4096 * 32768),
    ("rowwise", 0): ("LN fwd chain: LN(y)*g+x -> LN(.)*g2, 263168x512", 0, 2 * M * d * 4),
    ("rowwise", 1): ("LN bwd (+add), 263168x512", 0, 2 * M * d * 4),
    ("rowwise", 2): ("GEGLU+LN fwd, 263168 x (4096 -> 2048)", 0, 2 * M * (4096 + 2048)),
    ("rowwise", 3): ("GEGLU+LN bwd", 0, 2 * M * (2048 + 4096 + 4096)),
}
SCALE = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3,
         "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def col(hdr, units, row, name):
    i = hdr.index(name)
    return float(row[i].replace(",", "")) * SCALE.get(units[i], 1.0)


def main():
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    out = ["# ncu `--set full` captures, round 1 (B200, `tools/profile_kernels.py` shapes = cfg2 text tower)",
           "",
           "Source: `profiles/%s_ncu_{gemm,attn,nce,rowwise}_{raw,details}.csv` (exported on the GPU box with" % TAG,
           "`ncu -i … --page raw|details --csv`; captured with `--clock-control none`, one kernel at a time, so",
           "times are cold-cache/serialised and clocks are higher than inside a full step).",
           "Peaks (MEASURED_PEAKS.json): HBM %s GB/s, bf16 %s TFLOP/s burst / %s sustained." %
           (peaks.get("hbm_gbs"), peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")),
           "",
           "| kernel | launch | time ms | DRAM traffic GB | algorithmic GB | TFLOP/s (alg.) | frac of bf16 burst peak | DRAM GB/s | frac of HBM peak | tensor-pipe active % | regs |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    g_traffic = g_alg = 0.0
    for name in ("gemm", "attn", "nce", "rowwise"):
        hdr, units, data = load(ROOT / "profiles" / f"{TAG}_ncu_{name}_raw.csv")
        tk = [c for c in hdr if "pipe_tensor_cycles_active" in c and "pct" in c]
        for i, row in enumerate(data):
            if not row or not row[0].strip().isdigit():
                continue
            kn = row[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            dur = col(hdr, units, row, "gpu__time_duration.sum")
            traffic = col(hdr, units, row, "dram__bytes_read.sum") + col(hdr, units, row, "dram__bytes_write.sum")
            label, fl, by = ALG.get((name, i), (kn, 0, 0))
            tf = fl / dur / 1e9 if fl else 0.0
            gbs = traffic / dur * 1e3
            tens = row[hdr.index(tk[0])] if tk else ""
            regs = row[hdr.index("launch__registers_per_thread")]
            out.append(f"| `{kn}` | {label} | {dur:.3f} | {traffic:.3f} | {by / 1e9:.3f} | "
                       f"{tf:.0f} | {tf / peaks.get('bf16_tflops', 1640.6):.2f} | {gbs:.0f} | "
                       f"{gbs / peaks.get('hbm_gbs', 6485.5):.2f} | {tens} | {regs} |")
            if name == "gemm":
                g_traffic += traffic
                g_alg += by / 1e9
    out += ["",
            f"GEMM DRAM traffic / algorithmic bytes over the four launches: {g_traffic / g_alg:.3f} "
            "(no wasted re-reads; operands are streamed once, the weight tile is L2 resident)."]
    (ROOT / "profiles" / f"{TAG}_ncu_summary.md").write_text("\n".join(out) + "\n")
    (ROOT / "profiles" / "gemm_traffic.json").write_text(json.dumps(
        {"traffic_over_algorithmic": round(g_traffic / g_alg, 4),
         "source": f"profiles/{TAG}_ncu_gemm_raw.csv: QKV fwd, FF-up fwd/dgrad/wgrad at cfg2 text shapes"}, indent=1))
    print("\n".join(out))


if __name__ == "__main__":
    main()
