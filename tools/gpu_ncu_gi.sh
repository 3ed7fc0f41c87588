#!/bin/bash
# ncu capture of the GI film kernel (one launch, full set + source)
mkdir -p gpurun_out
TAG=${1:-gi}
ncu --set full --clock-control none --import-source on -k regex:k_trace_camera -s 2 -c 1 -f -o gpurun_out/gi_$TAG python tools/ncu_target_gi.py > gpurun_out/ncu_gi_$TAG.log 2>&1
tail -3 gpurun_out/ncu_gi_$TAG.log
ncu -i gpurun_out/gi_$TAG.ncu-rep --page raw --csv > gpurun_out/gi_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/gi_$TAG.ncu-rep --page source --csv > gpurun_out/gi_${TAG}_source.csv 2>/dev/null
rm -f gpurun_out/gi_$TAG.ncu-rep
ls -la gpurun_out/ | tail -4
