#!/bin/bash
# GPU call 2: parity after the packed-fp32x2 FF epilogues, FF / attention micro-benchmarks with the
# tuning switches, ncu captures (source-level stalls) of ff_bwd and the short attention kernels,
# micro-batch size sweep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c2_pytest.log 2>&1
tail -3 gpurun_out/c2_pytest.log
{
echo "== ff_bench prefetch on"; timeout 200 python tools/ff_bench.py 50176,768 39936,512
echo "== ff_bench prefetch off"; XCLIP_TOOLS_TUNE="0=0" timeout 200 python tools/ff_bench.py 50176,768 39936,512
for c in 4 3 2 1; do echo "== attn_bench fwd CTAs/SM $c"; XCLIP_TOOLS_TUNE="1=$c" timeout 200 python tools/attn_bench.py 512,98,12 512,78,8; done
} > gpurun_out/c2_micro.log 2>&1
cat gpurun_out/c2_micro.log
A="python tools/attn_bench.py"
tools/ncu_kernel.sh ff_bwd gemm_pair_kernelILi0ELi1ELi3E 2 1 python tools/ff_bench.py 50176,768
tools/ncu_kernel.sh ff_up gemm_pair_kernelILi0ELi0ELi1E 2 1 python tools/ff_bench.py 50176,768
tools/ncu_kernel.sh attn_small_fwd attn_fwd_small_kernelILi128ELb0E 2 1 $A 512,98,12
tools/ncu_kernel.sh attn_small_bwd attn_bwd_small_kernelILb0E 2 1 $A 512,98,12
for n in ff_bwd ff_up attn_small_fwd attn_small_bwd; do
  python tools/ncu_stalls.py gpurun_out/prof_${n}_cuda.csv "$n (cuda source view)" > gpurun_out/prof_${n}_stalls.md 2>/dev/null
  rm -f gpurun_out/prof_${n}_raw.csv
done
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-eager-baseline --no-e2e"
for cfg in "512 auto" "1024 0" "1024 auto" "768 auto" "512 0"; do
  set -- $cfg
  ( timeout 300 $B --microbatch $1 --retain $2 ) > gpurun_out/c2_bench_mb$1_r$2.json 2> gpurun_out/c2_bench_mb$1_r$2.err
done
for f in gpurun_out/c2_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["config"].get("step_plan"), d["config"].get("peak_hbm_bytes_allocated"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
python -c "import torch; print('mem_get_info', torch.cuda.mem_get_info(0))"
