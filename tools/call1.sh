#!/bin/bash
# GPU call 1 of the resumed round-2 session: full GPU test suite, timelines, retention A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export XCLIP_BENCH_VERBOSE=1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c1_pytest.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-eager-baseline --no-e2e"
( timeout 300 $B --retain 0 ) > gpurun_out/c1_bench_retain0.json 2> gpurun_out/c1_bench_retain0.err
( timeout 300 $B --retain auto ) > gpurun_out/c1_bench_auto.json 2> gpurun_out/c1_bench_auto.err
( timeout 300 $B --retain auto --microbatch 256 ) > gpurun_out/c1_bench_auto_mb256.json 2> gpurun_out/c1_bench_auto_mb256.err
( timeout 300 python tools/timeline.py --retain auto --out gpurun_out/timeline_auto ) > gpurun_out/c1_timeline.log 2>&1
tail -3 gpurun_out/c1_pytest.log
for f in gpurun_out/c1_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["config"].get("step_plan"), d["config"].get("peak_hbm_bytes_allocated"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
head -8 gpurun_out/timeline_auto.md
