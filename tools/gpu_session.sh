#!/bin/bash
# One gpurun call: parity tests, build probes (SAT8 on/off) with tree checksums, launch list of one build.
mkdir -p gpurun_out
{
echo "== pytest default"; python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== build sat8=1"; python tools/probe_build.py 2>&1 | tail -3
echo "== build sat8=0"; VRT_BUILD_SAT8=0 python tools/probe_build.py 2>&1 | tail -3
echo "== build sorted"; VRT_BUILD_SORTED=1 python tools/probe_build.py 2>&1 | tail -3
} > gpurun_out/session.log 2>&1
tail -40 gpurun_out/session.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_build_r2c.csv python tools/probe_build_one.py > gpurun_out/ncu_build.log 2>&1
tail -2 gpurun_out/ncu_build.log
