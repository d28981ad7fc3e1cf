#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/variant_probe.py --config chime5 torchain_b200/libtorchain_b200.so > $O/r02_release_probe.txt 2>&1
timeout 600 python tools/variant_probe.py --no-parity --config chime5_tied torchain_b200/libtorchain_b200.so >> $O/r02_release_probe.txt 2>&1
cat $O/r02_release_probe.txt
TORCHAIN_HOST_TRACE=1 timeout 90 python tools/host_trace_probe.py > $O/r02_host_trace_release.txt 2>&1
tail -10 $O/r02_host_trace_release.txt
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -m gpu -x -k "host" > $O/r02_release_tests.log 2>&1
tail -2 $O/r02_release_tests.log
