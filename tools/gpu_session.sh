#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest default"; python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in "1" "0"; do
  echo "== tri64=$v"
  VRT_TRI64=$v timeout 300 python tools/probe_quick.py 11 2>&1 | tail -2
done
} > gpurun_out/session.log 2>&1
tail -40 gpurun_out/session.log
