#!/bin/bash
# exp(y) tile gathered through L1 (plain loads) in the forward (fe), backward (be), both (fbe) against ld.global.cg (default)
set -u
O=gpurun_out/r02_etile_l1_probe.txt
: > $O
for cfgB in "chime5 0" "chime5_tied 0" "chime5 64" "large_tree 0"; do
  set -- $cfgB
  echo "== config $1 batch $2: default, fe, be, fbe, default" >> $O
  timeout 900 python tools/variant_probe.py --config $1 --batch $2 --steps 20 "" build/variants/fe.so build/variants/be.so build/variants/fbe.so "" >> $O 2>&1
done
cat $O
