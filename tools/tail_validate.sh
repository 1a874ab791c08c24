#!/bin/bash
# GPU validation of the opt-in attention paths.  usage: tools/tail_validate.sh [stage ...]
#   stages: tail variants bwd16 elect bench   (default: all, in that order)
# Every step runs under a SHORT timeout: a device-side hang (a hardware barrier or a full-mask
# shuffle has no timeout of its own) then costs seconds, not the round's GPU budget.
cd "$(dirname "$0")/.."
stages=${*:-tail variants bwd16 elect bench}
summ() { grep -E "passed|failed|error|Timeout|FAIL|time|ALL OK|SOME" | tail -${1:-6}; }

for st in $stages; do
case $st in
tail)
  echo "== [tail] attention tests with XCLIP_ATTN_TAIL=1"
  XCLIP_ATTN_TAIL=1 timeout 60 python -m pytest tests/test_gpu_attention.py -q -m gpu -x 2>&1 | summ 4
  echo "== [tail] model parity with XCLIP_ATTN_TAIL=1"
  XCLIP_ATTN_TAIL=1 timeout 90 python -m pytest tests/test_gpu_clip_parity.py -q -m gpu -x 2>&1 | summ 3
  XCLIP_ATTN_TAIL=1 timeout 45 python tools/pp_check.py 2>&1 | summ 8
  ;;
variants)
  for v in 0 1 2 3 4 5 6 7 8 9; do
    echo "== [variants] forward ping-pong variant $v"
    XCLIP_ATTN_PP_VARIANT=$v timeout 45 python tools/pp_check.py 2>&1 | summ 8
  done
  ;;
bwd16)
  echo "== [bwd16] 16-warp backward kernel"
  XCLIP_ATTN_BWD16=1 timeout 60 python -m pytest tests/test_gpu_attention.py -q -m gpu -x -k bwd 2>&1 | summ 3
  XCLIP_ATTN_BWD16=1 timeout 45 python tools/pp_check.py 2>&1 | grep -E "bwd time|rror"
  XCLIP_ATTN_BWD16=1 XCLIP_ATTN_TAIL=1 timeout 60 python -m pytest tests/test_gpu_attention.py -q -m gpu -x -k bwd 2>&1 | summ 3
  ;;
elect)
  if [ -f x_clip_b200/libxclip_b200_elect.so ]; then
    echo "== [elect] elect.sync build (made with XCLIP_BUILD_ELECT=1 python -m x_clip_b200.build before gpurun)"
    XCLIP_LIB_VARIANT=elect timeout 240 python -m pytest tests -q -m gpu -x 2>&1 | summ 3
    XCLIP_LIB_VARIANT=elect timeout 60 python tools/pp_check.py 2>&1 | summ 8
    XCLIP_LIB_VARIANT=elect timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
  else
    echo "== [elect] libxclip_b200_elect.so not built - skipped"
  fi
  ;;
bench)
  echo "== [bench] default, then with the tail path"
  timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
  XCLIP_ATTN_TAIL=1 timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
  ;;
esac
done
