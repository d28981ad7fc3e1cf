#!/bin/bash
# small batches: where does a frame go?  fixed cost (deg 0 = 2 arcs per state) vs slope per arc, with CTA traces
set -u
O=gpurun_out/r02_smallbatch_timeline.txt
: > $O
for B in 32 64 128; do
  for deg in 0 7 14 28; do
    echo "== B=$B deg=$deg" >> $O
    TL_TRACE=1 timeout 300 python tools/den_timeline.py --batch $B --deg $deg >> $O 2>&1
  done
done
cat $O
