#!/bin/bash
# usage: tools/ncu_kernel.sh <name> <mangled-name-regex> <skip> <count> <cmd...>
# Full ncu capture of `count` launches (after skipping `skip`) of the kernels whose MANGLED name
# matches the regex; keeps compact CSV pages (details, raw, per-source-line stalls) in gpurun_out/.
name=$1; regex=$2; skip=$3; cnt=$4; shift 4
rep=/tmp/prof_$name
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
  -k regex:$regex -s $skip -c $cnt -f -o $rep "$@" > gpurun_out/prof_$name.log 2>&1
ncu -i $rep.ncu-rep --page details --csv > gpurun_out/prof_${name}_details.csv 2>/dev/null
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --print-source cuda --csv > gpurun_out/prof_${name}_cuda.csv 2>/dev/null
sz=$(stat -c %s $rep.ncu-rep 2>/dev/null || echo 0)
echo "$name: rep $sz bytes"
