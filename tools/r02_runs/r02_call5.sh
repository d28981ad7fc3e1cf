#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -rA -m gpu -k "tied" > $O/r02_tied_tests.log 2>&1
grep -E "^parity|passed|failed|^FAILED|^ERROR|Error" $O/r02_tied_tests.log | tail -25
timeout 300 python tools/variant_probe.py --config chime5_tied torchain_b200/libtorchain_b200.so > $O/r02_tied_probe.txt 2>&1
TORCHAIN_TIED=0 timeout 300 python tools/variant_probe.py --config chime5_tied --no-parity torchain_b200/libtorchain_b200.so >> $O/r02_tied_probe.txt 2>&1
timeout 300 python tools/variant_probe.py --config chime5 --no-parity torchain_b200/libtorchain_b200.so >> $O/r02_tied_probe.txt 2>&1
cat $O/r02_tied_probe.txt
