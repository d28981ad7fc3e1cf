#!/bin/bash
# final capture of round 2 after the normaliser-load fix: GPU suite, bench, reference arm, ncu launch list, ncu full of
# the per-arc and tied kernels, smoke
set -u
bash tools/r02_capture.sh
bash tools/r02_runs/r02_tied_ncu.sh
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_final.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
