#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c11_pytest.log 2>&1
tail -4 gpurun_out/c11_pytest.log
( timeout 200 python tools/ff_bench.py 50176,768 39936,512 ) > gpurun_out/c11_ff_bench.log 2>&1
cat gpurun_out/c11_ff_bench.log
( timeout 200 python tools/ln_bench.py ) > gpurun_out/c11_ln_bench.log 2>&1
cat gpurun_out/c11_ln_bench.log
( timeout 300 python bench.py --steps 8 --warmup 4 --no-extras --no-cpu-baseline --no-eager-baseline ) > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('/root/repo/gpurun_out/c11_bench.json').read().strip().splitlines()[-1])
    print("BENCH", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["step_plan"], d["clocks"])
    kf=d['kernel_families']
    print({k:(v['ms'],v.get('tflops')) for k,v in kf.items() if k in ('ff_up','ff_down','ff_bwd','cast','gemm_fwd')})
except Exception as e:
    print("BENCH ERR", e)
PY
