#!/bin/bash
# The multi-GPU points of BASELINE.json configs 3-5 on ONE box with 8 GPUs (one JSON line per point into
# gpurun_out/r01_multi_gpu_points.jsonl), plus the box's host<->device copy floor at 1/2/4/8 ranks.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/scale_points.sh'
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r01_multi_gpu_points.jsonl
: > $OUT
PORT=29511
run() {  # n_gpus, then bench.py arguments
  local n=$1; shift
  PORT=$((PORT + 1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $n --steps 10 --warmup 3 "$@" 2>>gpurun_out/scale_points.err | tail -1 >> $OUT
}
nproc > gpurun_out/box8_nproc.txt; nvidia-smi topo -m > gpurun_out/box8_topo.txt 2>&1
# config 3: cfg-2 graph, 512 sequences sharded over 2 / 4 / 8 GPUs
run 2 --config chime5 --batch 256
run 4 --config chime5 --batch 128
run 8 --config chime5 --batch 64
# config 4: large tree, 256 sequences per GPU (weak) and 256 in total (32 per GPU)
run 8 --config large_tree --batch 256
run 8 --config large_tree --batch 32
# config 5: batch / length sweep at 8 GPUs (total sequences 64 .. 2048)
run 8 --config chime5 --batch 8 --frames 50
run 8 --config chime5 --batch 256 --frames 50
run 8 --config chime5 --batch 256 --frames 200
run 8 --config chime5 --batch 256 --frames 500
# host<->device copy floor
for n in 1 2 4 8; do
  PORT=$((PORT + 1))
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
    tools/pcie_probe.py 2>>gpurun_out/scale_points.err | tail -1 >> gpurun_out/r01_pcie_floor.jsonl
done
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print('bad line', l[:80]); continue
    print(d['n_gpus'], d['config']['workload'][:70], '%.3f ms' % d['ms_per_step'], '%.2f Mfr/s' % (d['value'] / 1e6), 'e2e %.2f ms' % d['e2e']['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
"
cat gpurun_out/r01_pcie_floor.jsonl
