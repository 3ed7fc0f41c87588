#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest default"; python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for s in "" _mb8; do
  echo "== variant '$s'"
  VRT_LIB_SUFFIX=$s timeout 300 python tools/probe_quick.py 11 2>&1 | tail -2
done
echo "== e2e build"; python tools/probe_build_e2e.py 2>&1 | tail -4
} > gpurun_out/session.log 2>&1
tail -40 gpurun_out/session.log
