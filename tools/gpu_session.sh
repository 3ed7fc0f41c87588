#!/bin/bash
mkdir -p gpurun_out
{
for s in "" _sn; do echo "== variant '$s'"; VRT_LIB_SUFFIX=$s python tools/probe_quick.py 11 2>&1 | tail -1 | cut -c1-200; done
} > gpurun_out/session.log 2>&1
tail -30 gpurun_out/session.log
