#!/bin/bash
# per-state vs per-frame fixed cost: N sweep at 2 arcs per state, and the same arc count with half the states
set -u
O=gpurun_out/r02_state_cost.txt
: > $O
for N in 2500 5000 10000 20000 40000; do
  echo "== B=128 deg=0 N=$N" >> $O
  TL_TRACE=1 timeout 300 python tools/den_timeline.py --batch 128 --deg 0 --states $N 2>&1 | grep -E "^A=|fwd frames|per-frame" >> $O
done
echo "== B=128 deg=58 N=5000 (same arcs as cfg 2, half the states)" >> $O
TL_TRACE=1 timeout 300 python tools/den_timeline.py --batch 128 --deg 58 --states 5000 2>&1 | grep -E "^A=|fwd frames|per-frame" >> $O
echo "== B=128 deg=13 N=20000 (same arcs as cfg 2, twice the states)" >> $O
TL_TRACE=1 timeout 300 python tools/den_timeline.py --batch 128 --deg 13 --states 20000 2>&1 | grep -E "^A=|fwd frames|per-frame" >> $O
cat $O
