"""Host<->device copy floor of the box at N ranks (why the host-buffer entry's e2e number stops
scaling with the GPU count): every rank copies `--mb` MB in and the same out of pinned memory,
H2D and D2H concurrently on two streams, all ranks at once (barrier-aligned), timed with CUDA
events, max over ranks.  Launch like bench.py:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py
"""
import argparse
import json
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=74.9568)  # config 2: 6400 rows x 2928 pdfs x 4 B
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(args.mb * 1e6 / 4)
h_in, h_out = torch.empty(n, pin_memory=True).normal_(), torch.empty(n, pin_memory=True)
d_in, d_out = torch.empty(n, device=dev), torch.randn(n, device=dev)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(h2d, d2h):
    def once():
        if h2d:
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
    for _ in range(3):
        once()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_in.wait_event(e0)
    s_out.wait_event(e0)
    for _ in range(args.iters):
        once()
    torch.cuda.current_stream(dev).wait_stream(s_in)
    torch.cuda.current_stream(dev).wait_stream(s_out)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


out = {"n_gpus": world, "mb_each_way": args.mb, "host_cpus": os.cpu_count()}
for name, (a, b) in {"h2d": (True, False), "d2h": (False, True), "both": (True, True)}.items():
    ms = run(a, b)
    out[name + "_ms"] = round(ms, 3)
    out[name + "_gbs_per_gpu_per_dir"] = round(args.mb / ms, 1)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
