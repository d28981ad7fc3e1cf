#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --no-parity --config chime5 $V/ff4.so $V/ff43r.so $V/ff8.so $V/f4nofast.so > $O/r02_fast_probe2.txt 2>&1
timeout 900 python tools/variant_probe.py --no-parity --config large_tree $V/ff4.so torchain_b200/libtorchain_b200.so >> $O/r02_fast_probe2.txt 2>&1
timeout 900 python tools/variant_probe.py --no-parity --config chime5 --batch 256 $V/ff4.so torchain_b200/libtorchain_b200.so >> $O/r02_fast_probe2.txt 2>&1
timeout 900 python tools/variant_probe.py --no-parity --config chime5 --batch 64 $V/ff4.so torchain_b200/libtorchain_b200.so >> $O/r02_fast_probe2.txt 2>&1
cat $O/r02_fast_probe2.txt
