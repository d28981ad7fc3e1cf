#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python tools/variant_probe.py --config chime5 torchain_b200/libtorchain_b200.so > $O/r02_ldcg_probe.txt 2>&1
timeout 900 python tools/variant_probe.py --no-parity --config chime5_tied torchain_b200/libtorchain_b200.so >> $O/r02_ldcg_probe.txt 2>&1
timeout 900 python tools/variant_probe.py --no-parity --config chime5 --batch 256 torchain_b200/libtorchain_b200.so >> $O/r02_ldcg_probe.txt 2>&1
cat $O/r02_ldcg_probe.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "host or odd" 2>&1 | tail -2
