#!/bin/bash
# usage: tools/probe_variants.sh suffix1 suffix2 ...   (perf + output checksums of prebuilt libvrt<suffix>.so variants)
for s in "$@"; do
  echo "== variant '$s'"
  VRT_LIB_SUFFIX=$s timeout 300 python tools/probe_quick.py 11 2>&1 | tail -2
done
