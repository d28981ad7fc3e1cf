"""Batch/length sweep of the chain loss on one GPU (BASELINE.json configs 2, 4 and the
per-GPU points of config 5).  Prints one JSON line per point:
    python tools/sweep.py [--points "chime5:128:50,chime5:256:50,large_tree:256:50"]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from torchain_b200 import _lib, io
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="chime5:8:50,chime5:32:50,chime5:64:50,chime5:128:50,chime5:256:50,"
                                         "chime5:128:100,chime5:128:200,chime5:256:500,large_tree:32:50,large_tree:256:50")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    lib = _lib.load()
    sys.path.insert(0, os.path.join(ROOT, "tools", "microbench"))
    import mb as _mb
    mb = _mb.load()
    graphs = {}
    l2_peak = mb.tb_microbench_l2_gather(0, 32 << 20, 512, 400)
    for pt in args.points.split(","):
        name, B, T = pt.split(":")
        B, T = int(B), int(T)
        c = bench.make_workload(name, B, T, 0)
        den, sup = c["den"], c["sup"]
        P = c["P"]
        if name not in graphs:
            graphs[name] = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost,
                                                           den.start, den.final_cost, den.num_pdfs)
        g = graphs[name]
        s = io.Supervision.from_arrays(sup.weight, B, T, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                                       sup.arc_cost, sup.final_cost)
        y = torch.randn(T * B, P, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 2.0
        d = torch.empty_like(y)
        res = np.zeros(3, dtype=np.float32)
        stream = torch.cuda.current_stream()

        def step():
            ok = lib.torchain_ComputeChainObjfAndDeriv_ex(
                g.ptr, s.ptr, y.data_ptr(), T * B, P, P, res.ctypes.data_as(ctypes.c_void_p), None, d.data_ptr(), P,
                None, 0, c["l2"], c["leaky"], 0.0, -1.0, 0, ctypes.c_void_p(stream.cuda_stream))
            assert ok == 1, _lib.last_error()

        for _ in range(4):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        hbm_alg, gather_alg = bench.alg_bytes(c)
        print(json.dumps({"config": name, "B": B, "T": T, "N": c["N"], "P": P, "A": den.num_arcs, "ms_per_step": ms,
                          "frames_per_s": B * T / ms * 1e3, "gather_TBps": gather_alg * B * T / ms / 1e9,
                          "gather_frac_of_measured_l2_peak": gather_alg * B * T / ms / 1e6 / l2_peak,
                          "hbm_GBps": hbm_alg * B * T / ms / 1e6, "objf_per_frame": float(res[0] / res[2])}), flush=True)
        del s, y, d


if __name__ == "__main__":
    main()
