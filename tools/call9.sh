#!/bin/bash
# profiling call: ncu --set full per kernel, ncu launch list of one reduced step, final bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/r2_profile.sh 2>&1 | tail -14
for n in attn_small_fwd attn_small_bwd ff_bwd ff_up pair_wgrad; do
  cp /tmp/prof_$n.ncu-rep gpurun_out/ 2>/dev/null
done
rm -f gpurun_out/prof_*_cuda.csv
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_step.csv python tools/one_step.py 1536 768 1 > gpurun_out/c9_one_step.log 2>&1
tail -2 gpurun_out/c9_one_step.log; wc -l gpurun_out/launches_step.csv
( XCLIP_BENCH_VERBOSE=1 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/c9_bench_driver_like.json 2> gpurun_out/c9_bench_driver_like.err
grep "bench rank 0" gpurun_out/c9_bench_driver_like.err | cut -c1-200 | tail -6
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c9_bench_driver_like.json').read().strip().splitlines()[-1])
    print("FULL", d["value"], d["ms_per_step"], d["e2e"], d["config"]["step_plan"], d["clocks"])
except Exception as e:
    print("FULL ERR", e)
PY
( timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/c9_bench_reference_arm.json 2> gpurun_out/c9_bench_reference_arm.err
cut -c1-400 gpurun_out/c9_bench_reference_arm.json
