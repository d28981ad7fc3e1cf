#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 600 python tools/variant_probe.py --no-parity --config chime5_tied $V/noaux.so > $O/r02_noaux_probe.txt 2>&1
timeout 600 python tools/variant_probe.py --no-parity --config chime5 $V/noaux.so >> $O/r02_noaux_probe.txt 2>&1
cat $O/r02_noaux_probe.txt
