#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/variant_probe.py --config chime5_tied torchain_b200/libtorchain_b200.so > $O/r02_tied_probe4.txt 2>&1
cat $O/r02_tied_probe4.txt
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -rA -m gpu -k "tied" > $O/r02_tied_tests2.log 2>&1
grep -E "^parity|passed|failed|^FAILED|^ERROR|Error" $O/r02_tied_tests2.log | tail -12
