#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --no-parity --config chime5_tied torchain_b200/libtorchain_b200.so $V/t42.so $V/t43.so $V/t41.so $V/t22.so > $O/r02_tied_sa_probe.txt 2>&1
cat $O/r02_tied_sa_probe.txt
