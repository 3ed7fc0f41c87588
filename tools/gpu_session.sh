#!/bin/bash
# One gpurun call: parity tests, perf + checksums of prebuilt library variants, a short bench line.
mkdir -p gpurun_out
{
echo "== pytest default"; python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== pytest VRT_TRACE_WS=1"; VRT_TRACE_WS=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for s in "" _tp _tp6; do
  echo "== variant '$s'"
  VRT_LIB_SUFFIX=$s timeout 300 python tools/probe_quick.py 11 2>&1 | tail -2
done
echo "== bench"; python bench.py --steps 10 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python tools/show_bench.py gpurun_out/bench_quick.json
} > gpurun_out/session.log 2>&1
tail -60 gpurun_out/session.log
