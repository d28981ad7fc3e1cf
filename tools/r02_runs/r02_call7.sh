#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 600 python tools/variant_probe.py --config chime5_tied torchain_b200/libtorchain_b200.so $V/tf128.so $V/tf32.so $V/tdbg.so > $O/r02_tied_probe3.txt 2>&1
cat $O/r02_tied_probe3.txt
