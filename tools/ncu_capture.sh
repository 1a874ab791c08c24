#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel-regex> <count> <cmd...>
# Full ncu capture of a few launches of one kernel; only compact CSV pages are kept in
# gpurun_out/ (the .ncu-rep with sources is kept only when small).
name=$1; regex=$2; cnt=$3; shift 3
rep=/tmp/prof_$name
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$regex -c $cnt -f -o $rep "$@" > gpurun_out/prof_$name.log 2>&1
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
ncu -i $rep.ncu-rep --page details --csv > gpurun_out/prof_${name}_details.csv 2>/dev/null
# per-instruction and per-source-line stall samples (summarise with tools/ncu_stalls.py)
ncu -i $rep.ncu-rep --page source --csv > gpurun_out/prof_${name}_sass.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --print-source cuda --csv > gpurun_out/prof_${name}_cuda.csv 2>/dev/null
sz=$(stat -c %s $rep.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -lt 12000000 ]; then cp $rep.ncu-rep gpurun_out/; fi
echo "$name: rep $sz bytes"
