"""Small chain-loss calls for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Covers V = 1, 2, 4 kernels, a split (heavy-tailed) graph and the host-buffer entry."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCHAIN_SPLIT_CAP", "9")

import torch  # noqa: E402
from torchain_b200 import _lib, io, synth  # noqa: E402

lib = _lib.load()
for (B, T, N, P, hubs) in [(5, 4, 60, 17, 0), (40, 3, 80, 21, 1), (70, 3, 90, 33, 0)]:
    den = synth.make_den_fst(B, N, P, mean_deg=3.0)
    if hubs:
        den = synth.add_hubs(den, hubs)
    sup = synth.make_supervision(B + 1, den, B, T, num_paths=2)
    y = synth.make_nnet_output(B + 2, B, T, P)
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    s = io.Supervision.from_arrays(sup.weight, B, T, P, sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost,
                                   sup.final_cost)
    yt = torch.from_numpy(y).cuda()
    d = torch.empty_like(yt)
    x = torch.empty_like(yt)
    res = np.zeros(3, dtype=np.float32)
    ok = lib.my_lib_ComputeChainObjfAndDeriv(g.ptr, s.ptr, yt.data_ptr(), B * T, P, P, res.ctypes.data_as(ctypes.c_void_p),
                                             d.data_ptr(), P, x.data_ptr(), P, 0.01, 0.1, 0.1, 0, None)
    assert ok == 1, _lib.last_error()
    torch.cuda.synchronize()
    dh = np.empty_like(y)
    ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, y.ctypes.data_as(ctypes.c_void_p), B * T, P, P,
                                                    res.ctypes.data_as(ctypes.c_void_p), dh.ctypes.data_as(ctypes.c_void_p),
                                                    P, None, 0, 0.01, 0.1, 0.0, 0)
    assert ok == 1, _lib.last_error()
    print("ok", B, T, N, P, res, float(np.abs(dh - d.cpu().numpy()).max()))

# strided (frame, sequence, pdf) entry on a (B, P, T) tensor, relayout kernels, merged supervision
for (B, T, N, P) in [(5, 4, 60, 17), (37, 6, 80, 33)]:
    den = synth.make_den_fst(B + 7, N, P, mean_deg=3.0)
    parts = [synth.make_supervision(100 + i, den, 1, T, num_paths=2) for i in range(B)]
    hs = [io.Supervision.from_arrays(p.weight, 1, T, P, p.arc_begin, p.arc_dst, p.arc_label, p.arc_cost, p.final_cost)
          for p in parts]
    s = io.Supervision.merge(hs)
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    y2 = torch.from_numpy(synth.make_nnet_output(B + 2, B, T, P)).cuda()
    y3 = y2.view(T, B, P).permute(1, 2, 0).contiguous()
    d3, x3 = torch.empty_like(y3), torch.empty_like(y3)
    st3 = (ctypes.c_int64 * 3)(1, P * T, T)
    res = np.zeros(3, dtype=np.float32)
    ok = lib.torchain_ComputeChainObjfAndDeriv_strided(g.ptr, s.ptr, y3.data_ptr(), T, B, P, st3,
                                                       res.ctypes.data_as(ctypes.c_void_p), None, d3.data_ptr(), st3,
                                                       x3.data_ptr(), st3, 0.01, 0.1, 0.1, 1.0, 0, None)
    assert ok == 1, _lib.last_error()
    rows = torch.empty(T * B, P + 1, device="cuda")
    assert lib.torchain_to2d(y3.data_ptr(), B, P, T, rows.data_ptr(), P + 1, 0, None) == 1, _lib.last_error()
    back = torch.empty_like(y3)
    assert lib.torchain_from2d(rows.data_ptr(), P + 1, B, P, T, back.data_ptr(), 1.0, 0, None) == 1, _lib.last_error()
    torch.cuda.synchronize()
    d2 = torch.empty_like(y2)
    res2 = np.zeros(3, dtype=np.float32)
    ok = lib.my_lib_ComputeChainObjfAndDeriv(g.ptr, s.ptr, y2.data_ptr(), B * T, P, P, res2.ctypes.data_as(ctypes.c_void_p),
                                             d2.data_ptr(), P, None, 0, 0.01, 0.1, 0.0, 0, None)
    assert ok == 1, _lib.last_error()
    torch.cuda.synchronize()
    print("ok strided/relayout/merge", B, T, N, P, res, res2, bool(torch.equal(back, y3)), bool(torch.equal(rows[:, :P], y2)),
          float((d3.permute(2, 0, 1).reshape(T * B, P) - d2).abs().max()))

# round 2: the run-based (tied) kernels -- device and host-buffer entries, a remainder pass on the per-arc kernels --
# the TMA relayout kernels (aligned shapes), a prefetched supervision whose pooled block is reused, chain-egs parsing
os.environ["TORCHAIN_TIED"] = "1"   # (the split cap of 9 set above would keep these graphs on the per-arc kernels)
for (B, T, N, P) in [(70, 4, 90, 24), (130, 3, 60, 16)]:
    den = synth.make_den_fst_tied(B, N, P, mean_deg=6.0)
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    assert g.tied, (g.n_run, den.num_arcs)
    y = synth.make_nnet_output(B + 2, B, T, P)
    yt = torch.from_numpy(y).cuda()
    side = torch.cuda.Stream()
    for rep in range(3):   # three supervisions through the same pooled block
        sup = synth.make_supervision(B + 10 + rep, den, B, T, num_paths=2)
        s = io.Supervision.from_arrays(sup.weight, B, T, P, sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost,
                                       sup.final_cost).prefetch(0, side)
        d = torch.empty_like(yt)
        res = np.zeros(3, dtype=np.float32)
        ok = lib.my_lib_ComputeChainObjfAndDeriv(g.ptr, s.ptr, yt.data_ptr(), B * T, P, P, res.ctypes.data_as(ctypes.c_void_p),
                                                 d.data_ptr(), P, None, 0, 0.01, 0.1, 0.0, 0, None)
        assert ok == 1, _lib.last_error()
        torch.cuda.synchronize()
        dh = np.empty_like(y)
        ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, y.ctypes.data_as(ctypes.c_void_p), B * T, P, P,
                                                        res.ctypes.data_as(ctypes.c_void_p), dh.ctypes.data_as(ctypes.c_void_p),
                                                        P, None, 0, 0.01, 0.1, 0.0, 0)
        assert ok == 1, _lib.last_error()
        del s
    print("ok tied", B, T, N, P, res, float(np.abs(dh - d.cpu().numpy()).max()))
del os.environ["TORCHAIN_TIED"]
for (B, P, T) in [(5, 64, 12), (3, 48, 50), (2, 36, 4)]:
    x = torch.randn(B, P, T, device="cuda")
    rows = torch.empty(T * B, P + 4, device="cuda")
    assert lib.torchain_to2d(x.data_ptr(), B, P, T, rows.data_ptr(), P + 4, 0, None) == 1, _lib.last_error()
    back = torch.empty_like(x)
    assert lib.torchain_from2d(rows.data_ptr(), P + 4, B, P, T, back.data_ptr(), 1.0, 0, None) == 1, _lib.last_error()
    torch.cuda.synchronize()
    print("ok TMA relayout", B, P, T, bool(torch.equal(back, x)),
          bool(torch.equal(rows[:, :P], x.permute(2, 0, 1).reshape(T * B, P))))
lib.torchain_supervision_pool_trim()
