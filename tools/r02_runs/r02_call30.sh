#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --config chime5_tied torchain_b200/libtorchain_b200.so $V/nov4.so > $O/r02_tied_auxv4_probe.txt 2>&1
cat $O/r02_tied_auxv4_probe.txt
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x -k "tied" 2>&1 | tail -2
