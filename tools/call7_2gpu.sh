#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/c7_pytest_multirank.log 2>&1
tail -3 gpurun_out/c7_pytest_multirank.log
( XCLIP_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-extras ) > gpurun_out/c7_bench_n2.json 2> gpurun_out/c7_bench_n2.err
grep "bench rank 0" gpurun_out/c7_bench_n2.err | cut -c1-300 | tail -12
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c7_bench_n2.json').read().strip().splitlines()[-1])
    print("N2", d["value"], d["ms_per_step"], d["e2e"], d["config"]["step_plan"], d["multirank_parity"], d["with_grad_sync"])
except Exception as e:
    print("N2 ERR", e)
PY
