#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --no-parity --config chime5 torchain_b200/libtorchain_b200.so $V/b120.so $V/b112.so $V/b104.so $V/f120.so $V/f104.so > $O/r02_maxnreg_probe.txt 2>&1
cat $O/r02_maxnreg_probe.txt
