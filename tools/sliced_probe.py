"""Feasibility probe for the sequence-sliced recurrence: microseconds per frame of
tb_microbench_sliced (tools/microbench) on a synthetic graph of config 2's shape (N=10000, 30 arcs per state, P=2928, 128 sequences),
to hold against the current kernels' 25.9 us (forward) / 42.6 us (backward) per frame."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "microbench"))
import mb as _mb  # noqa: E402

mb = _mb.load()
names = {0: "forward-like (exp(y) slice in smem)", 1: "backward-like, fp32 shared atomics (ATOMS.CAST.SPIN)",
         2: "backward-like, fixed-point shared atomics (ATOMS.ADD)"}
for S, mode in ((16, 0), (8, 0), (8, 1), (8, 2)):
    for depth in (4, 8):
        us = mb.tb_microbench_sliced(0, S, depth, mode, 50)
        us1 = mb.tb_microbench_sliced(0, S, depth, mode + 4, 50)  # exp(y) staged once: the gather phase alone
        print("S=%2d sequences/CTA  depth=%d  %-55s %8.2f us/frame  (%7.2f without per-frame staging)" % (
            S, depth, names[mode], us, us1), flush=True)
