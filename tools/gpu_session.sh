#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== build"; timeout 300 python tools/probe_build.py 2>&1 | tail -3
echo "== build sorted"; VRT_BUILD_SORTED=1 timeout 300 python tools/probe_build.py 2>&1 | tail -3
} > gpurun_out/session.log 2>&1
tail -30 gpurun_out/session.log
