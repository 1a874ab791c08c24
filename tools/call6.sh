#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c6_pytest.log 2>&1
tail -3 gpurun_out/c6_pytest.log
( XCLIP_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 8 --warmup 4 ) > gpurun_out/c6_bench_full.json 2> gpurun_out/c6_bench_full.err
grep -v "^Traceback\|^  " gpurun_out/c6_bench_full.err | cut -c1-250 | tail -20
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c6_bench_full.json').read().strip().splitlines()[-1])
    print("FULL", d["value"], d["ms_per_step"], d["e2e"], d["config"]["step_plan"], d["config"]["peak_hbm_bytes_allocated"], d["clocks"])
    print({k:(v['ms'],v.get('tflops'),v.get('gbs')) for k,v in d['kernel_families'].items()})
except Exception as e:
    print("FULL ERR", e)
PY
( timeout 300 python tools/timeline.py --microbatch 768 --out gpurun_out/timeline_mb768 ) > gpurun_out/c6_timeline.log 2>&1
head -8 gpurun_out/timeline_mb768.md
