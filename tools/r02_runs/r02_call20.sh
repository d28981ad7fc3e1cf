#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
L=torchain_b200/libtorchain_b200.so
( timeout 600 python tools/variant_probe.py --config chime5_tied $L
  TORCHAIN_KEEP_ETILE=0 timeout 600 python tools/variant_probe.py --no-parity --config chime5_tied $L
  timeout 600 python tools/variant_probe.py --no-parity --config chime5 $L
  TORCHAIN_KEEP_ETILE=1 timeout 600 python tools/variant_probe.py --config chime5 $L
  TORCHAIN_KEEP_ETILE=1 timeout 600 python tools/variant_probe.py --no-parity --config chime5 --batch 256 $L
  timeout 600 python tools/variant_probe.py --no-parity --config chime5 --batch 256 $L ) > $O/r02_keep_etile_probe.txt 2>&1
cat $O/r02_keep_etile_probe.txt
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -m gpu -x -k "tied or host_buffer or cfg2_full_size" > $O/r02_keep_tests.log 2>&1
tail -3 $O/r02_keep_tests.log
