#!/bin/bash
# normaliser rows read with L1-allocating loads (default now) against ld.global.cg (build/variants/s_cg.so)
set -u
O=gpurun_out/r02_norm_probe.txt
: > $O
for cfgB in "chime5 0" "chime5 64" "chime5 32" "chime5 256" "chime5_tied 0" "large_tree 0"; do
  set -- $cfgB
  echo "== config $1 batch $2 (first line: L1-allocating loads; second: ld.global.cg)" >> $O
  timeout 600 python tools/variant_probe.py --config $1 --batch $2 --steps 20 "" build/variants/s_cg.so >> $O 2>&1
done
for deg in 0 28; do
  echo "== timeline B=128 deg=$deg" >> $O
  TL_TRACE=1 timeout 300 python tools/den_timeline.py --batch 128 --deg $deg 2>&1 | grep -E "^A=|fwd frames|per-frame" >> $O
done
cat $O
