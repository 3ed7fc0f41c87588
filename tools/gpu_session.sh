#!/bin/bash
# One gpurun call: parity tests for the kernel variants, then perf + checksums of prebuilt library variants.
mkdir -p gpurun_out
{
echo "== pytest default (warp-synchronous + hull)"; python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== pytest VRT_TRACE_WS=0 (per-ray + hull)"; VRT_TRACE_WS=0 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== pytest VRT_TRACE_WS=0 VRT_HULL=0 (round-1 kernel)"; VRT_TRACE_WS=0 VRT_HULL=0 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for s in "" _mb6 _mb5; do for ws in 0 1; do for hull in 0 1; do
  echo "== variant '$s' ws=$ws hull=$hull"
  VRT_LIB_SUFFIX=$s VRT_TRACE_WS=$ws VRT_HULL=$hull timeout 300 python tools/probe_quick.py 11 2>&1 | tail -2
done; done; done
} > gpurun_out/session.log 2>&1
tail -60 gpurun_out/session.log
