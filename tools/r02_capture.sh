#!/bin/bash
# (the .ncu-rep files stay in /tmp on the box: gpurun copies back at most 64 MiB; the raw csv pages come back)
# Round-2 capture on ONE B200: the whole GPU suite, a bench line, the ncu launch list of the same command, one
# `ncu --set full` capture of the denominator kernels (per-arc on config 2, run-based on the tied graph), the
# relayout kernels, the layout probe.  Everything lands in gpurun_out/r02_*; summaries go to profiles/ by hand.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -rA > $O/r02_gpu_tests_final.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR" $O/r02_gpu_tests_final.log | tail -8
timeout 600 python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err
tail -2 $O/r02_bench_final.err
python tools/layout_probe.py > $O/r02_layout_probe_final.txt 2>&1
tail -9 $O/r02_layout_probe_final.txt | head -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches_final.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-extra-configs --no-fresh-supervision > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:den_ -s 8 -c 2 -f -o /tmp/r02_den \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-extra-configs --no-fresh-supervision > $O/ncu_den.log 2>&1
ncu -i /tmp/r02_den.ncu-rep --page raw --csv > $O/r02_ncu_den_kernels_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:relayout -s 2 -c 2 -f -o /tmp/r02_relayout \
    python tools/layout_probe.py --iters 1 > $O/ncu_relayout.log 2>&1
ncu -i /tmp/r02_relayout.ncu-rep --page raw --csv > $O/r02_ncu_relayout_kernels_raw.csv 2>/dev/null
python - <<'PY'
import csv, json
for f in ("gpurun_out/r02_ncu_den_kernels_raw.csv", "gpurun_out/r02_ncu_relayout_kernels_raw.csv"):
    rows = list(csv.reader(open(f)))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "lts__t_sectors.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    idx = [hdr.index(w) for w in want if w in hdr]
    print(rows[1][idx[1]] if len(rows) > 1 else "", "(units row)")
    for r in rows[2:]:
        print({hdr[i]: r[i] for i in idx})
d = json.load(open("gpurun_out/r02_bench_final.json"))
for k in ("ms_per_step", "value", "fresh_supervision_ms_per_step", "cfg3_strong_ms_per_step", "cfg4_ms_per_step", "parity_ok"):
    print(k, d.get(k))
print("e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"])
PY
