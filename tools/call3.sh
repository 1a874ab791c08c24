#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c3_pytest.log 2>&1
tail -3 gpurun_out/c3_pytest.log
{
echo "== attn_bench default (table-free fast path)"; timeout 200 python tools/attn_bench.py 512,98,12 512,78,8 1024,32,8
echo "== attn_bench + L2 prefetch of the next item"; XCLIP_TOOLS_TUNE="2=1" timeout 200 python tools/attn_bench.py 512,98,12 512,78,8 1024,32,8
} > gpurun_out/c3_micro.log 2>&1
cat gpurun_out/c3_micro.log
A="python tools/attn_bench.py"
tools/ncu_kernel.sh ff_bwd gemm_pair_kernelILi0ELi1ELi3E 2 1 python tools/ff_bench.py 50176,768
tools/ncu_kernel.sh attn_small_fwd attn_fwd_small_kernelILi128ELb0E 2 1 $A 512,98,12
tools/ncu_kernel.sh attn_small_bwd attn_bwd_small_kernelILb0E 2 1 $A 512,98,12
for n in ff_bwd attn_small_fwd attn_small_bwd; do
  cp /tmp/prof_$n.ncu-rep gpurun_out/ 2>/dev/null
  rm -f gpurun_out/prof_${n}_raw.csv gpurun_out/prof_${n}_cuda.csv
done
( timeout 400 python tools/ab_step.py "mb=512,retain=auto" "mb=768,retain=auto" "mb=640,retain=auto" "mb=512,retain=auto,accum=0" "mb=1024,retain=auto" --rounds=2 --steps=3 ) > gpurun_out/c3_ab.log 2>&1
tail -8 gpurun_out/c3_ab.log
( timeout 300 python tools/ab_step.py "mb=512,retain=auto" "mb=768,retain=auto" --rounds=2 --steps=3 --alloc=expandable ) > gpurun_out/c3_ab_expandable.log 2>&1
tail -4 gpurun_out/c3_ab_expandable.log
