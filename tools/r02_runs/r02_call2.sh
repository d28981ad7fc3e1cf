#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rA > $O/r02_gpu_tests2.log 2>&1
grep -E "^parity|passed|failed|^FAILED|^ERROR" $O/r02_gpu_tests2.log | tail -25
timeout 400 python bench.py > $O/r02_bench2.json 2> $O/r02_bench2.err
tail -3 $O/r02_bench2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench2.json'))
for k in ("ms_per_step", "value", "fresh_supervision_ms_per_step", "fresh_supervision_serial_ms_per_step", "cfg3_strong_ms_per_step",
          "cfg4_ms_per_step", "parity_ok", "parity_deriv_rel_err"):
    print(k, d.get(k))
print("e2e", d["e2e"]["ms_per_step"], "launch", d["roofline"]["launch_ms"], "frac", d["roofline"]["frac"], d["clocks"])
PY
TORCHAIN_HOST_TRACE=1 timeout 90 python tools/host_trace_probe.py 2>&1 | tail -12
