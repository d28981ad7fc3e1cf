#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --no-parity --config chime5_tied $V/t41.so $V/t81.so > $O/r02_tied_sa_probe2.txt 2>&1
timeout 900 python tools/variant_probe.py --no-parity --config chime5 torchain_b200/libtorchain_b200.so $V/f41.so $V/t81.so >> $O/r02_tied_sa_probe2.txt 2>&1
cat $O/r02_tied_sa_probe2.txt
