#!/bin/bash
# final 1-GPU records of a round: bench line, reference arm, ncu full capture of the frame step, launch list of a bench run
TAG=${1:-r2n}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || tail -5 gpurun_out/bench_$TAG.err
python tools/show_bench.py gpurun_out/bench_$TAG.json | cut -c1-500
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
bash tools/gpu_ncu.sh $TAG > /dev/null 2>&1; rm -f gpurun_out/trace_$TAG.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gi --no-build-soup --build-reps 1 > gpurun_out/ncu_bench_$TAG.log 2>&1
tail -c 300 gpurun_out/ncu_bench_$TAG.log
