#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py $V/sync.so $V/a4.so $V/a8.so $V/afonly.so $V/t640.so $V/t768.so $V/t768a3.so > $O/r02_variants_async.txt 2>&1
cat $O/r02_variants_async.txt
