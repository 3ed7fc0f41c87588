#!/bin/bash
# final captures of the round: ncu of the shipped ray kernel + launch list of a bench run + the default bench line
mkdir -p gpurun_out
bash tools/gpu_ncu.sh r2i
rm -f gpurun_out/trace_r2i.ncu-rep
python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -2 gpurun_out/bench_r2i.err; python tools/show_bench.py gpurun_out/bench_r2i.json 2>/dev/null | head -3 | cut -c1-400
bash tools/gpu_launchlist.sh bench_r2i python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gi --no-build-soup
