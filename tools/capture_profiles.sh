#!/bin/bash
# End-of-round measurement pass on ONE B200 (gpurun --timeout 900 -- 'bash tools/capture_profiles.sh'):
# bench line, ncu launch list of the same command, ncu --set full of the denominator and relayout kernels,
# microbenchmarks, timeline, layout probe.  Everything lands in gpurun_out/; summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
O=gpurun_out
python bench.py > $O/r01_bench_final.json 2> $O/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r01_bench_reference_arm.json 2>> $O/bench_final.err
python tools/layout_probe.py > $O/r01_layout_probe.txt 2>&1
python tools/den_timeline.py > $O/r01_den_timeline_final.txt 2>&1
python - > $O/r01_microbench_final.txt 2>&1 <<'PY'
import sys
sys.path.insert(0, 'tools/microbench')
import mb as _mb
mb = _mb.load()
for rb in (512, 256, 128, 64, 32):
    print("gather", rb, mb.tb_microbench_l2_gather(0, 32 << 20, rb, 400))
    print("red", rb, mb.tb_microbench_l2_red(0, 32 << 20, rb, 400))
print("barrier us", mb.tb_microbench_grid_barrier(0, 2000))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_ncu_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:den_ -s 8 -c 2 -f -o $O/r01_den \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_den.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:relayout -s 2 -c 2 -f -o $O/r01_relayout \
    python tools/layout_probe.py --iters 1 > $O/ncu_relayout.log 2>&1
ls -la $O | tail -20
tail -c 400 $O/r01_bench_final.json
