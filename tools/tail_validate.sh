#!/bin/bash
# One-shot GPU validation of the opt-in attention paths (tail token, ping-pong variants).
# Every step runs under a SHORT timeout: a hang costs seconds, not the round's GPU budget.
cd "$(dirname "$0")/.."
echo "== attention tests, tail path on"
XCLIP_ATTN_TAIL=1 timeout 60 python -m pytest tests/test_gpu_attention.py -q -m gpu -x 2>&1 | tail -4
echo "== model parity, tail path on"
XCLIP_ATTN_TAIL=1 timeout 90 python -m pytest tests/test_gpu_clip_parity.py -q -m gpu -x 2>&1 | tail -3
for v in 0 1 2 3 4 5 6 7; do XCLIP_ATTN_PP_VARIANT=$v timeout 45 python tools/pp_check.py 2>&1 | grep -E "FAIL|time|ALL|SOME|rror"; done
XCLIP_ATTN_TAIL=1 timeout 45 python tools/pp_check.py 2>&1 | grep -E "FAIL|time|ALL|SOME|rror"
echo "== 16-warp backward kernel"
XCLIP_ATTN_BWD16=1 timeout 60 python -m pytest tests/test_gpu_attention.py -q -m gpu -x -k bwd 2>&1 | tail -3
XCLIP_ATTN_BWD16=1 timeout 45 python tools/pp_check.py 2>&1 | grep -E "bwd time|rror"
XCLIP_ATTN_BWD16=1 XCLIP_ATTN_TAIL=1 timeout 60 python -m pytest tests/test_gpu_attention.py -q -m gpu -x -k bwd 2>&1 | tail -3
echo "== bench (default), then with the tail path"
timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
XCLIP_ATTN_TAIL=1 timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
if [ -f x_clip_b200/libxclip_b200_elect.so ]; then
  echo "== elect.sync build (XCLIP_BUILD_ELECT=1 python -m x_clip_b200.build, made before gpurun)"
  XCLIP_LIB_VARIANT=elect timeout 240 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
  XCLIP_LIB_VARIANT=elect timeout 60 python tools/pp_check.py 2>&1 | grep -E "FAIL|time|ALL|SOME|rror"
  XCLIP_LIB_VARIANT=elect timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
fi
