#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ff.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/c8_pytest_ff.log 2>&1
tail -2 gpurun_out/c8_pytest_ff.log
( timeout 200 python tools/attn_layout_probe.py ) > gpurun_out/c8_layout_probe.log 2>&1
cat gpurun_out/c8_layout_probe.log
( timeout 200 python tools/ff_bench.py 50176,768 75264,768 59904,512 ) > gpurun_out/c8_ff_bench.log 2>&1
cat gpurun_out/c8_ff_bench.log
