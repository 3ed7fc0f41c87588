#!/bin/bash
mkdir -p gpurun_out
{
for s in "" _hp; do
  echo "== variant '$s'"
  VRT_LIB_SUFFIX=$s timeout 300 python tools/probe_quick.py 11 2>&1 | tail -2
done
} > gpurun_out/session.log 2>&1
tail -40 gpurun_out/session.log
