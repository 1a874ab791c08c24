#!/bin/bash
# End-of-round validation on a GPU box (gpurun): full GPU test suite, smoke(), the driver's bench command.
#   gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'            (1 GPU)
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_validate.sh 2' (adds the 2-rank tests + torchrun bench)
cd "$(dirname "$0")/.."
N=${1:-1}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/validate_pytest.log 2>&1
tail -4 gpurun_out/validate_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
if [ "$N" = "1" ]; then
  ( XCLIP_BENCH_VERBOSE=1 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err
else
  ( XCLIP_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err
fi
grep "bench rank 0" gpurun_out/validate_bench.err | cut -c1-220 | tail -8
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/validate_bench.json').read().strip().splitlines()[-1])
    print("BENCH", d["n_gpus"], "GPUs:", d["value"], "pairs/s,", d["ms_per_step"], "ms/step; e2e", d["e2e"]["value"],
          "; plan", d["config"]["step_plan"], "; clocks", d["clocks"], "; parity", d.get("multirank_parity"))
except Exception as e:
    print("BENCH ERR", e)
PY
