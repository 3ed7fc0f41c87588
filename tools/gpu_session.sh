#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest"; python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench"; python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; tail -3 gpurun_out/bench_r2e.err; python tools/show_bench.py gpurun_out/bench_r2e.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_r2e.json').read().strip().splitlines()[-1]); print('gi', d.get('gi'))"
} > gpurun_out/session.log 2>&1
tail -40 gpurun_out/session.log
