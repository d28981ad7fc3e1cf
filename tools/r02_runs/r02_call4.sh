#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py $V/r512.so $V/r640.so $V/r640nf.so $V/r768.so $V/r896.so > $O/r02_variants_roles.txt 2>&1
cat $O/r02_variants_roles.txt
