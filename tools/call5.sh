#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c5_pytest.log 2>&1
tail -3 gpurun_out/c5_pytest.log
{
echo "== ff_bench variant 0 (ld.global epilogue)"; XCLIP_TOOLS_TUNE="0=0" timeout 200 python tools/ff_bench.py 50176,768 39936,512
echo "== ff_bench variant 1 (TMA-pipelined epilogue)"; XCLIP_TOOLS_TUNE="0=1" timeout 200 python tools/ff_bench.py 50176,768 39936,512
} > gpurun_out/c5_micro.log 2>&1
cat gpurun_out/c5_micro.log
XCLIP_TOOLS_TUNE="0=1" tools/ncu_kernel.sh ff_bwd2 gemm_pair_kernelILi0ELi1ELi4E 2 1 python tools/ff_bench.py 50176,768
cp /tmp/prof_ff_bwd2.ncu-rep gpurun_out/ 2>/dev/null; rm -f gpurun_out/prof_ff_bwd2_raw.csv gpurun_out/prof_ff_bwd2_cuda.csv
( timeout 400 python tools/ab_step.py "mb=768,retain=auto,tune=0:0" "mb=768,retain=auto,tune=0:1" --rounds=3 --steps=3 --alloc=expandable ) > gpurun_out/c5_ab.log 2>&1
tail -6 gpurun_out/c5_ab.log
( XCLIP_BENCH_VERBOSE=1 timeout 900 python bench.py --tune 0=1 ) > gpurun_out/c5_bench_full.json 2> gpurun_out/c5_bench_full.err
tail -12 gpurun_out/c5_bench_full.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c5_bench_full.json').read().strip().splitlines()[-1])
    print("FULL", d["value"], d["ms_per_step"], d["e2e"], d["config"]["step_plan"], d["clocks"])
except Exception as e:
    print("FULL ERR", e)
PY
