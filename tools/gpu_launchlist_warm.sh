#!/bin/bash
# launch list with warm caches (ncu --cache-control none): the kernels' busy time without the profiler's cache flushes
mkdir -p gpurun_out
TAG=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c ${NCU_MAX_LAUNCHES:-800} --csv --log-file gpurun_out/launches_$TAG.csv "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/ncu_$TAG.log
