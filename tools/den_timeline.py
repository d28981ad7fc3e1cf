"""Per-phase timeline of the persistent denominator kernel (globaltimer stamps).

    python tools/den_timeline.py [--config chime5] [--batch B] [--frames T]
Prints the mean forward / backward frame time and the slowest frames.
"""
import argparse
import ctypes
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
    ap.add_argument("--config", default="chime5")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--deg", type=float, default=-1.0, help="override the mean out-degree of the synthetic den graph")
    ap.add_argument("--states", type=int, default=0)
    args = ap.parse_args()
    from torchain_b200 import synth
    if args.deg >= 0:
        synth.CONFIGS[args.config]["mean_deg"] = args.deg
    if args.states:
        synth.CONFIGS[args.config]["N"] = args.states
    lib = _lib.load()
    c = bench.make_workload(args.config, args.batch, args.frames, 0)
    den, sup = c["den"], c["sup"]
    B, T, P = c["B"], c["T"], c["P"]
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    s = io.Supervision.from_arrays(sup.weight, B, T, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                                   sup.arc_cost, sup.final_cost)
    y = torch.randn(T * B, P, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 2.0
    d = torch.empty_like(y)
    res = np.zeros(3, dtype=np.float32)

    def step():
        ok = lib.my_lib_ComputeChainObjfAndDeriv(g.ptr, s.ptr, y.data_ptr(), T * B, P, P,
                                                 res.ctypes.data_as(ctypes.c_void_p), d.data_ptr(), P, None, 0,
                                                 c["l2"], c["leaky"], 0.0, 0,
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert ok == 1, _lib.last_error()

    for _ in range(int(os.environ.get("TL_WARMUP", "3"))):
        step()
    lib.torchain_profile_enable(2 if os.environ.get("TL_TRACE") else 1)
    step()
    torch.cuda.synchronize()
    n = 2 * T + 6
    tl = np.zeros(n, dtype=np.int64)
    got = lib.torchain_profile_timeline(g.ptr, tl.ctypes.data_as(ctypes.c_void_p), n)
    lib.torchain_profile_enable(0)
    assert got == n, got
    # forward launch: [start, prologue, T frames, end]; backward launch: [start, prologue, T frames, end]
    f, b = tl[:T + 3], tl[T + 3:]
    df, db = np.diff(f) / 1e3, np.diff(b) / 1e3
    fwd, bwd = df[1:1 + T], db[1:1 + T]
    print("A=%d N=%d" % (den.num_arcs, den.num_states))
    print("config %s B=%d T=%d: forward launch %.1f us, gap %.1f us (numerator), backward launch %.1f us, total %.1f us" %
          (args.config, B, T, (f[-1] - f[0]) / 1e3, (b[0] - f[-1]) / 1e3, (b[-1] - b[0]) / 1e3, (b[-1] - f[0]) / 1e3))
    print("prologue %.1f us | fwd frames mean %.2f min %.2f max %.2f (sum %.1f) | fwd tail %.1f | bwd prologue %.1f |"
          " bwd frames mean %.2f min %.2f max %.2f (sum %.1f) | epilogue %.1f" %
          (df[0], fwd.mean(), fwd.min(), fwd.max(), fwd.sum(), df[-1], db[0], bwd.mean(), bwd.min(), bwd.max(),
           bwd.sum(), db[-1]))
    if os.environ.get("TL_TRACE"):
        fr, gr = ctypes.c_int(0), ctypes.c_int(0)
        n = lib.torchain_profile_cta_trace(g.ptr, None, 0, ctypes.byref(fr), ctypes.byref(gr))
        tr = np.zeros(n, dtype=np.int64)
        lib.torchain_profile_cta_trace(g.ptr, tr.ctypes.data_as(ctypes.c_void_p), n, ctypes.byref(fr), ctypes.byref(gr))
        nw = n // (2 * fr.value * gr.value) - 4
        wtr = tr[2 * fr.value * gr.value * 4:].reshape(2, fr.value, gr.value, nw).astype(np.float64) / 1e3
        tr = tr[:2 * fr.value * gr.value * 4].reshape(2, fr.value, gr.value, 4).astype(np.float64) / 1e3
        for ph, name, frames in ((0, "fwd", range(2, T)), (1, "bwd", range(1, T - 1))):
            rows = []
            for t in frames:
                x = tr[ph, t]                                   # [grid][4]: min main, max main, max aux, barrier exit
                prev = tr[ph, t - 1 if ph == 0 else t + 1][:, 3]  # barrier exit of the previous step
                rows.append([np.median(x[:, 0] - prev), np.median(x[:, 1] - prev), (x[:, 1] - prev).max(),
                             np.median(x[:, 2] - prev), (x[:, 2] - prev).max(), np.median(x[:, 3] - prev)])
            # is the CTA-to-CTA spread systematic?  per-CTA mean busy time over frames vs its frame-to-frame std
            busy = np.stack([tr[ph, t][:, 1] - tr[ph, t - 1 if ph == 0 else t + 1][:, 3] for t in frames])  # [frames][grid]
            m, sd = busy.mean(axis=0), busy.std(axis=0)
            order = np.argsort(m)
            print("%s CTA busy time: mean over CTAs %.2f, min %.2f, max %.2f, std across CTAs %.2f, mean per-CTA std over frames %.2f" %
                  (name, m.mean(), m.min(), m.max(), m.std(), sd.mean()))
            print("   slowest CTAs:", [(int(c), round(float(m[c]), 1)) for c in order[-8:]], " fastest:",
                  [(int(c), round(float(m[c]), 1)) for c in order[:6]])
            if os.environ.get("TL_TRACE_DUMP"):
                np.save(os.environ["TL_TRACE_DUMP"] + "_%s.npy" % name, busy)
            # per-warp finish time relative to the CTA's barrier exit, averaged over frames and CTAs
            wfin = np.stack([wtr[ph, t] - tr[ph, t - 1 if ph == 0 else t + 1][:, 3:4] for t in frames])  # [frames][grid][warps]
            print("   mean finish time by warp index:", np.round(wfin.mean(axis=(0, 1)), 1).tolist())
            print("   std over frames of a (cta, warp) finish time: %.2f; std over (cta, warp) of the mean: %.2f" %
                  (wfin.std(axis=0).mean(), wfin.mean(axis=0).std()))
            r = np.array(rows).mean(axis=0)
            print("%s per-frame, us since the CTA left the previous barrier: first main warp done %.1f (median over CTAs) |"
                  " last main warp done %.1f median, %.1f max | last aux warp done %.1f median, %.1f max | barrier exit %.1f" %
                  (name, r[0], r[1], r[2], r[3], r[4], r[5]))
    A = den.num_arcs
    print("ideal @14.5 TB/s L2 gather: fwd frame %.2f us (A*B*8 B), bwd frame gathers %.2f us + red A*B*4 B" %
          (A * B * 8 / 14.5e6, A * B * 8 / 14.5e6))
    if os.environ.get("TL_VERBOSE"):
        print("fwd frames:", np.round(fwd, 1).tolist())
        print("bwd frames:", np.round(bwd, 1).tolist())


if __name__ == "__main__":
    main()
