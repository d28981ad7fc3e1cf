#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 600 python tools/variant_probe.py --config chime5_tied torchain_b200/libtorchain_b200.so $V/x576.so $V/x576b.so $V/x608.so > $O/r02_aux_probe.txt 2>&1
timeout 600 python tools/variant_probe.py --config chime5 torchain_b200/libtorchain_b200.so $V/x576.so $V/x576b.so $V/x608.so >> $O/r02_aux_probe.txt 2>&1
cat $O/r02_aux_probe.txt
