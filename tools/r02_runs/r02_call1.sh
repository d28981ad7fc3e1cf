#!/bin/bash
# GPU call 1 of round 2: full-size oracle parity first (the verdict's top item), then the rest of the suite,
# the host-entry variants and a bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/r02_nproc.txt
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -rA -m gpu > $O/r02_fullsize.log 2>&1
grep -E "^parity|passed|failed|Error" $O/r02_fullsize.log | head -40
timeout 400 python -m pytest tests -q -m gpu --deselect tests/test_gpu_fullsize.py > $O/r02_gpu_tests.log 2>&1
tail -3 $O/r02_gpu_tests.log
TORCHAIN_HOST_TRACE=1 timeout 90 python tools/host_trace_probe.py > $O/r02_host_trace.txt 2>&1
python - "$O/r02_host_trace.txt" <<'PY'
import re, sys, statistics
name, rows = None, {}
for line in open(sys.argv[1]):
    if line.startswith("== "):
        name = line[3:].strip()
    m = re.search(r"compute_done=(\d+) end=(\d+)", line)
    if m and name:
        rows.setdefault(name, []).append((int(m.group(1)), int(m.group(2))))
for k, v in rows.items():
    print("%-75s compute_done %6.0f  end %6.0f us (median of %d)" % (
        k, statistics.median(a for a, _ in v), statistics.median(b for _, b in v), len(v)))
PY
timeout 200 python bench.py > $O/r02_bench_start.json 2> $O/r02_bench_start.err
python -c "
import json; d=json.load(open('$O/r02_bench_start.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['e2e']['ms_per_step'], d['clocks'])"
timeout 300 python tools/gather_probe.py > $O/r02_gather_probe.txt 2>&1
cat $O/r02_gather_probe.txt
