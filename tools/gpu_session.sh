#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest"; python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for sa in 0.6 0.7 0.85 0.95; do echo "== VRT_HULL_SA=$sa"; VRT_HULL_SA=$sa python tools/probe_soup_frame.py 2>&1 | tail -2; done
} > gpurun_out/session.log 2>&1
tail -30 gpurun_out/session.log
