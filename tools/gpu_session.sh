#!/bin/bash
mkdir -p gpurun_out
{
for s in "" _hf; do echo "== variant '$s'"; VRT_LIB_SUFFIX=$s python tools/probe_soup_frame.py 2>&1 | tail -2; VRT_LIB_SUFFIX=$s python tools/probe_quick.py 11 2>&1 | tail -1 | cut -c1-220; done
} > gpurun_out/session.log 2>&1
tail -30 gpurun_out/session.log
