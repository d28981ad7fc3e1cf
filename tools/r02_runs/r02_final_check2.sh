#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -rA > $O/r02_gpu_tests_final2.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR" $O/r02_gpu_tests_final2.log | tail -5
timeout 600 python bench.py > $O/r02_bench_final2.json 2> $O/r02_bench_final2.err
tail -2 $O/r02_bench_final2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm2.json 2>> $O/r02_bench_final2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_final2.json'))
for k in ("ms_per_step", "value", "fresh_supervision_ms_per_step", "cfg3_strong_ms_per_step", "cfg4_ms_per_step", "parity_ok", "gpu_launches"):
    print(k, d.get(k))
print("e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"])
r = json.load(open('gpurun_out/r02_bench_reference_arm2.json'))
print("reference arm", r["value"], r["cpu_baseline"])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
