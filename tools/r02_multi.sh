#!/bin/bash
# Multi-GPU check on ONE box with N GPUs (gpurun --gpus N -- 'bash tools/r02_multi.sh N [sweep]'): the NCCL test of the
# distributed op (N >= 2), the bench line at N exactly as the driver launches it and, with `sweep`, the points of
# BASELINE.json config 5 (batch / length sweep) on the same box, one JSON line each into gpurun_out/r02_sweep_nN.jsonl.
set -u
N=${1:-2}
SWEEP=${2:-}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_distributed.py -q -m gpu -rA > $O/r02_nccl_tests_n$N.log 2>&1
grep -E "passed|failed|skipped|^FAILED|^ERROR|^PASSED|^SKIPPED" $O/r02_nccl_tests_n$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 \
    bench.py --gpus $N --steps 20 --warmup 3 > $O/r02_scale_n$N.json 2> $O/r02_scale_n$N.err
tail -3 $O/r02_scale_n$N.err
python - $N <<'PY'
import json, sys
d = json.load(open('gpurun_out/r02_scale_n%s.json' % sys.argv[1]))
for k in ("n_gpus", "ms_per_step", "value", "fresh_supervision_ms_per_step", "cfg3_strong_seqs_per_gpu", "cfg3_strong_ms_per_step", "cfg3_strong_frames_per_sec",
          "cfg4_ms_per_step", "cfg4_frames_per_sec", "cfg5_64seqs_ms_per_step", "cfg5_64seqs_frames_per_sec", "parity_ok", "parity_deriv_rel_err"):
    print(k, d.get(k))
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
PY
if [ -n "$SWEEP" ]; then
  OUT=$O/r02_sweep_n$N.jsonl
  : > $OUT
  PORT=29800
  for pt in "64 50" "128 100" "256 50" "256 200" "256 500" "32 500"; do
    set -- $pt
    PORT=$((PORT + 1))
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 10 --warmup 3 --batch $1 --frames $2 --no-parity --no-fresh-supervision --no-extra-configs \
      2>>$O/r02_sweep.err | tail -1 >> $OUT
  done
  python - $OUT <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: print('bad line', l[:80]); continue
    print(d['n_gpus'], d['config']['workload'][60:130], '%.3f ms' % d['ms_per_step'], '%.2f Mfr/s' % (d['value'] / 1e6), 'e2e %.2f ms' % d['e2e']['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
PY
fi
