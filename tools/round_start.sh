#!/bin/bash
# First GPU call of a round on ONE B200 (gpurun --timeout 600 -- 'bash tools/round_start.sh r02'):
# the GPU suite on the code as it stands (tests added without a GPU at the end of the previous round run here
# for the first time), the host-side timeline of the e2e step with every experiment variant of the host-buffer
# entry, and a bench line.  Everything lands in gpurun_out/<round>_*.
set -u
R=${1:-r02}
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -q -x -m gpu > $O/${R}_gpu_tests.log 2>&1
tail -3 $O/${R}_gpu_tests.log
TORCHAIN_HOST_TRACE=1 timeout 60 python tools/host_trace_probe.py > $O/${R}_host_trace.txt 2>&1
grep -E "^==|objf" $O/${R}_host_trace.txt | head -20
python - "$O/${R}_host_trace.txt" <<'PY'
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
timeout 200 python bench.py > $O/${R}_bench_start.json 2> $O/${R}_bench_start.err
python -c "
import json; d=json.load(open('$O/${R}_bench_start.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['e2e']['ms_per_step'], d['clocks'])"
