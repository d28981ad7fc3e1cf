#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --no-parity --config chime5_tied $V/dbg2.so $V/dbg4.so $V/a4.so $V/a1.so > $O/r02_tied_phaseA_probe.txt 2>&1
cat $O/r02_tied_phaseA_probe.txt
