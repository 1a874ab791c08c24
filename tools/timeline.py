#!/usr/bin/env python
"""GPU timeline of ONE benchmark step (torch.profiler / CUPTI; no nsys in this image).

    python tools/timeline.py [--batch 4096] [--microbatch 512] [--retain auto] [--out gpurun_out/timeline]

Writes <out>.md: span of the step on the GPU, busy time (union of kernel intervals), idle time, the
kernels ranked by total time (ours AND torch's eager kernels / memsets / memcpys, which the C-ABI
profiler of bench.py does not see) and the largest idle gaps with the kernels on either side.
Diagnostic only - never a bench value (profiler attached)."""
from __future__ import annotations

import argparse
import collections
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--microbatch", type=int, default=512)
    ap.add_argument("--retain", default="auto")
    ap.add_argument("--out", default="gpurun_out/timeline")
    a = ap.parse_args()
    retain = a.retain if a.retain == "auto" else int(a.retain)

    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    run = bench.Runner(bench.WORKLOADS[a.workload][0], "nce", a.batch, a.microbatch, 0.5, dev, 0, 1, retain=retain)
    for _ in range(3):
        run.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run.step()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            t0 = e.time_range.start
            ev.append((t0, t0 + e.time_range.elapsed_us(), e.name))
    ev.sort()
    span = ev[-1][1] - ev[0][0]
    busy, cur_s, cur_e = 0.0, ev[0][0], ev[0][1]
    gaps = []
    prev_name = ev[0][2]
    for s, e, nm in ev[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, prev_name, nm))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
        prev_name = nm
    busy += cur_e - cur_s
    by = collections.defaultdict(lambda: [0, 0.0])
    for s, e, nm in ev:
        by[nm][0] += 1
        by[nm][1] += e - s
    gap_by = collections.defaultdict(lambda: [0, 0.0])
    for g, p, n in gaps:
        key = (p[:60], n[:60])
        gap_by[key][0] += 1
        gap_by[key][1] += g
    out = Path(a.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    lines = [f"# GPU timeline of one step ({a.workload}, {a.batch} pairs, micro-batch {a.microbatch}, retain {a.retain})",
             "", f"plan: {getattr(run.clip, 'last_step_plan', None)}", "",
             f"* span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms ({busy / span:.3f}), idle {(span - busy) / 1e3:.2f} ms",
             f"* {len(ev)} device activities, {len(gaps)} idle gaps "
             f"({sum(1 for g in gaps if g[0] > 20)} longer than 20 us, {sum(g[0] for g in gaps if g[0] > 20) / 1e3:.2f} ms in those)",
             "", "| kernel / activity | calls | total ms | share of span |", "|---|---|---|---|"]
    for nm, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append(f"| `{nm[:90]}` | {c} | {t / 1e3:.3f} | {t / span:.4f} |")
    lines += ["", "| idle gap between (previous -> next) | count | total ms |", "|---|---|---|"]
    for (p, n), (c, t) in sorted(gap_by.items(), key=lambda kv: -kv[1][1])[:30]:
        lines.append(f"| `{p}` -> `{n}` | {c} | {t / 1e3:.3f} |")
    out.with_suffix(".md").write_text("\n".join(lines) + "\n")
    out.with_suffix(".json").write_text(json.dumps(
        {"span_us": span, "busy_us": busy, "n": len(ev),
         "kernels": {k: v for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:80]}}))
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
