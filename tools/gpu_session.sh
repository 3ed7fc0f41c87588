#!/bin/bash
mkdir -p gpurun_out
{
python tools/probe_quick.py 11 2>&1 | tail -1 | cut -c1-200
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} > gpurun_out/session.log 2>&1
tail -30 gpurun_out/session.log
