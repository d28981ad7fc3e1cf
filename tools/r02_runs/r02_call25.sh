#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --no-parity --config chime5_tied torchain_b200/libtorchain_b200.so $V/b41.so $V/b22.so $V/b21.so $V/b43.so > $O/r02_tied_phaseB_probe.txt 2>&1
cat $O/r02_tied_phaseB_probe.txt
