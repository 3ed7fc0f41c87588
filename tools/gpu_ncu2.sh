#!/bin/bash
# two ncu captures in one call: the default (per-ray + hull) kernel and the warp-synchronous variant
bash tools/gpu_ncu.sh r2d
VRT_TRACE_WS=1 bash tools/gpu_ncu.sh r2d_ws
rm -f gpurun_out/trace_r2d_ws.ncu-rep
