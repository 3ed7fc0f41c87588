#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest gi + mirror + dropin"; timeout 900 python -m pytest tests/test_gpu_gi.py tests/test_gpu_cpp_mirror.py tests/test_gpu_dropin.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -6
echo "== gi probe"; timeout 300 python tools/probe_gi.py 2>&1 | tail -2
} > gpurun_out/session.log 2>&1
tail -30 gpurun_out/session.log
