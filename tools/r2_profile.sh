#!/bin/bash
# Round-2 ncu captures (one launch each, `--set full --clock-control none`) of the kernels the north
# star names + the GEMM family; CSV pages land in gpurun_out/, tools/ncu_r2_summary.py turns them into
# profiles/r2_ncu_summary.md.  Run on the GPU box:  tools/r2_profile.sh
cd "$(dirname "$0")/.."
A="python tools/attn_bench.py"
tools/ncu_kernel.sh attn_small_fwd attn_fwd_small_kernelILi128ELb0E 2 1 $A 512,98,12
tools/ncu_kernel.sh attn_small_bwd attn_bwd_small_kernelILb0E 2 1 $A 512,98,12
tools/ncu_kernel.sh attn_wg_fwd attn_fwd_wg_kernel 2 1 $A 1024,257,8
tools/ncu_kernel.sh attn_big_bwd 15attn_bwd_kernelE 2 1 $A 1024,257,8
tools/ncu_kernel.sh nce_fwd gemm_bf16_kernelILi256ELi0ELi0ELi1E 1 1 python tools/profile_kernels.py nce 3
tools/ncu_kernel.sh nce_bwd gemm_bf16_kernelILi256ELi0ELi0ELi2E 1 1 python tools/profile_kernels.py nce 3
tools/ncu_kernel.sh pair_store gemm_pair_kernelILi0ELi0ELi0E 3 1 python tools/ff_bench.py 50176,768
tools/ncu_kernel.sh ff_up gemm_pair_kernelILi0ELi0ELi1E 2 1 python tools/ff_bench.py 50176,768
tools/ncu_kernel.sh ff_down gemm_pair_kernelILi0ELi0ELi2E 2 1 python tools/ff_bench.py 50176,768
tools/ncu_kernel.sh ff_bwd gemm_pair_kernelILi0ELi1ELi4E 2 1 python tools/ff_bench.py 50176,768
tools/ncu_kernel.sh pair_wgrad gemm_pair_kernelILi1ELi1ELi0E 2 1 python tools/ff_bench.py 50176,768
tools/ncu_kernel.sh pair_dgrad gemm_pair_kernelILi0ELi1ELi0E 2 1 python tools/ff_bench.py 50176,768
