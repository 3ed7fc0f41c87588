#!/bin/bash
# usage: tools/probe_variants.sh suffix1 suffix2 ...   (perf of prebuilt libvrt<suffix>.so variants)
for s in "$@"; do
  echo "== variant '$s'"
  VRT_LIB_SUFFIX=$s timeout 300 python tools/probe_perf.py 11 2>&1 | grep -E "x4 hit16|x4 film|x4 frame"
done
