#!/bin/bash
mkdir -p gpurun_out
for h in 1 0; do
  VRT_HULL=$h python tools/ncu_target_soup.py > gpurun_out/plain_soup_$h.log 2>&1 || { tail gpurun_out/plain_soup_$h.log; exit 1; }
  VRT_HULL=$h ncu --set full --clock-control none -k regex:k_trace_camera -s 1 -c 1 -f -o gpurun_out/soup_h$h python tools/ncu_target_soup.py > gpurun_out/ncu_soup_$h.log 2>&1
  ncu -i gpurun_out/soup_h$h.ncu-rep --page raw --csv > gpurun_out/soup_h${h}_raw.csv 2>/dev/null
  rm -f gpurun_out/soup_h$h.ncu-rep
done
ls gpurun_out | grep soup_h
