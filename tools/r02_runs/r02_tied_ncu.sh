#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:den_tied -s 8 -c 2 -f -o /tmp/r02_den_tied_final python tools/variant_probe.py --child --config chime5_tied --no-parity --steps 2 > $O/ncu_tied.log 2>&1
ncu -i /tmp/r02_den_tied_final.ncu-rep --page raw --csv > $O/r02_ncu_den_tied_final_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r02_ncu_den_tied_final_raw.csv')))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "lts__t_sectors.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]: r[i] for i in idx})
PY
