#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
V=build/variants
timeout 900 python tools/variant_probe.py --no-parity --config chime5 torchain_b200/libtorchain_b200.so $V/bkeep.so $V/bbatch.so $V/bkeep22.so $V/bbatchfast.so > $O/r02_bwd_lds_probe.txt 2>&1
cat $O/r02_bwd_lds_probe.txt
