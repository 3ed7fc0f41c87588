#!/bin/bash
# launch list (ncu gpu__time_duration) of one command:  tools/gpu_launchlist.sh TAG cmd...
mkdir -p gpurun_out
TAG=$1; shift
"$@" > gpurun_out/plain_$TAG.log 2>&1 || { tail gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_MAX_LAUNCHES:-600} --csv --log-file gpurun_out/launches_$TAG.csv "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
