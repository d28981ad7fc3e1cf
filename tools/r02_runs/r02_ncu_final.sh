#!/bin/bash
# ncu launch list + --set full of the denominator launches of the final build (no test suite)
set -u
mkdir -p gpurun_out
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches_final.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-extra-configs --no-fresh-supervision > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:den_ -s 8 -c 2 -f -o /tmp/r02_den \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-extra-configs --no-fresh-supervision > $O/ncu_den.log 2>&1
ncu -i /tmp/r02_den.ncu-rep --page raw --csv > $O/r02_ncu_den_kernels_raw.csv 2>/dev/null
python tools/ncu_traffic_json.py $O/r02_ncu_den_kernels_raw.csv | head -30
