#!/bin/bash
mkdir -p gpurun_out
{
echo "== build mailbox"; timeout 300 python tools/probe_build.py 2>&1 | tail -3
echo "== build copy+sync"; VRT_BUILD_MAILBOX=0 timeout 300 python tools/probe_build.py 2>&1 | tail -3
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
} > gpurun_out/session.log 2>&1
tail -30 gpurun_out/session.log
