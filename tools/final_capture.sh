set -u
O=gpurun_out
python bench.py > $O/r01_bench_final.json 2> $O/bench_final.err
python tools/den_timeline.py > $O/r01_den_timeline_final.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_ncu_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:den_ -s 8 -c 2 -f -o $O/r01_den \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_den.log 2>&1
python -c "
import json; d=json.load(open('$O/r01_bench_final.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['e2e']['ms_per_step'], d['clocks'])"
