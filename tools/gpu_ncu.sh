#!/bin/bash
# ncu capture of the headline frame step (one launch, full set + source), per-ray + hull kernel by default
mkdir -p gpurun_out
TAG=${1:-r2a}
python tools/ncu_target.py 3 > gpurun_out/ncu_target_$TAG.log 2>&1 || { tail gpurun_out/ncu_target_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_trace_camera -s 2 -c 1 -f -o gpurun_out/trace_$TAG python tools/ncu_target.py 3 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ncu -i gpurun_out/trace_$TAG.ncu-rep --page raw --csv > gpurun_out/trace_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/trace_$TAG.ncu-rep --page source --csv > gpurun_out/trace_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
