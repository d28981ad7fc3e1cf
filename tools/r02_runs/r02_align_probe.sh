#!/bin/bash
# small batches: aligned streams (TORCHAIN_ALIGN_STREAMS) x split cap, B = 64 / 32 / 8 on the cfg-2 graph
set -u
O=gpurun_out/r02_align_probe.txt
: > $O
for B in 64 32; do
  for al in 0 1; do
    for cap in 192 24 16 12 8; do
      echo "== B=$B align=$al split_cap=$cap" >> $O
      TORCHAIN_ALIGN_STREAMS=$al TORCHAIN_SPLIT_CAP=$cap timeout 300 python tools/variant_probe.py --batch $B --steps 20 >> $O 2>&1
    done
  done
done
echo "== B=128 default" >> $O
timeout 300 python tools/variant_probe.py --steps 20 --no-parity >> $O 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "small_batches or b512 or ragged" 2>&1 | tail -3 >> $O
cat $O
