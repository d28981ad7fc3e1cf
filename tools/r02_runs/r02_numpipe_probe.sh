#!/bin/bash
# numerator occupancy chain pipelined behind the aux tile jobs (default) against the chain at the end (build/variants/num0.so)
set -u
O=gpurun_out/r02_numpipe_probe.txt
: > $O
for cfgB in "chime5 0" "chime5_tied 0" "chime5 64" "large_tree 0"; do
  set -- $cfgB
  echo "== config $1 batch $2 (first line: pipelined; second: chain at the end)" >> $O
  timeout 600 python tools/variant_probe.py --config $1 --batch $2 --steps 20 "" build/variants/num0.so "" build/variants/num0.so >> $O 2>&1
done
echo "== timeline B=128 deg=28" >> $O
TL_TRACE=1 timeout 300 python tools/den_timeline.py --batch 128 --deg 28 2>&1 | grep -E "^A=|fwd frames|per-frame" >> $O
cat $O
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
