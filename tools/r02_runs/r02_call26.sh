#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --config chime5_tied torchain_b200/libtorchain_b200.so > $O/r02_tied_final_probe.txt 2>&1
timeout 900 python tools/variant_probe.py --no-parity --config chime5_tied $V/b81.so $V/f81b81.so >> $O/r02_tied_final_probe.txt 2>&1
TORCHAIN_TIED=0 timeout 900 python tools/variant_probe.py --no-parity --config chime5_tied torchain_b200/libtorchain_b200.so >> $O/r02_tied_final_probe.txt 2>&1
cat $O/r02_tied_final_probe.txt
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x -k "tied" > $O/r02_tied_tests3.log 2>&1
tail -2 $O/r02_tied_tests3.log
