#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --config chime5 torchain_b200/libtorchain_b200.so $V/nolams.so > $O/r02_nolams_probe.txt 2>&1
cat $O/r02_nolams_probe.txt
