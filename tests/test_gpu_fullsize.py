"""GPU suite, full-size parity: the configurations the headline numbers are quoted on
(BASELINE.json cfg 2 / cfg 3 shard sizes / cfg 4) compared with the fp64 oracle on identical
inputs, through the C ABI, at the stated tolerances (objf, l2_term 1e-5 relative; derivative
1e-4 of the largest reference magnitude).  These are the cases the small-graph tests cannot
reach: N = 10 000 / 20 000 states, ~30 in-arcs per state summed in fp32 where Kaldi uses a
double, T = 50 chained per-frame normalisers, the shared-memory arc slices at their real size,
two contexts per launch (B = 256), the half / quarter-warp lane groups (B = 64 / 32) and a graph
whose arc slices do NOT fit shared memory.

The oracle runs on all host cores (fp64, OpenMP over states): ~5 s per cfg-2 minibatch on the
16-core GPU box, ~22 s on 8 cores.
"""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib as ol
from common import DERIV_RTOL, OBJF_RTOL, deriv_err, rel
from test_gpu_parity import handles, run_abi
from torchain_b200 import _lib, io, synth

pytestmark = pytest.mark.gpu


def _cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


_den_cache = {}


def _den(name):
    """The config's den graph with its oracle and library handles (built once per session)."""
    if name not in _den_cache:
        c = synth.CONFIGS[name]
        den = (synth.make_den_fst_tied if c.get("tied") else synth.make_den_fst)(1234, c["N"], c["P"],
                                                                                  mean_deg=c["mean_deg"])
        g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                            den.final_cost, den.num_pdfs)
        _den_cache[name] = (den, ol.OracleDen(den), g)
    return _den_cache[name]


def _compare(den, oden, g, B, T, l2, leaky, seed, weight=1.0, y_scale=2.0, want_xent=False, entry=run_abi):
    sup = synth.make_supervision(seed, den, B, T, num_paths=4, weight=weight)
    y = synth.make_nnet_output(seed + 1, B, T, den.num_pdfs, scale=y_scale)
    r = ol.chain(oden, ol.OracleSup(sup), y, l2=l2, leaky=leaky, mode="fp64", threads=_cores(), want_xent=want_xent)
    assert r["ok"]
    s = io.Supervision.from_arrays(sup.weight, B, T, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                                   sup.arc_cost, sup.final_cost)
    assert s.num_chains == B
    res, d, x = entry(g, s, y, l2, leaky, xent=0.1 if want_xent else 0.0, want_xent=want_xent)
    errs = dict(objf=rel(res[0], r["objf"]), l2=rel(res[1], r["l2_term"]) if l2 else 0.0,
                deriv=deriv_err(d, r["deriv"]))
    print("parity B=%d T=%d N=%d A=%d: objf %.9g vs %.9g (rel %.2e)  l2 rel %.2e  deriv %.2e" % (
        B, T, den.num_states, den.num_arcs, res[0], r["objf"], errs["objf"], errs["l2"], errs["deriv"]))
    assert errs["objf"] < OBJF_RTOL
    assert errs["l2"] < OBJF_RTOL
    assert res[2] == pytest.approx(r["weight"])
    assert errs["deriv"] < DERIV_RTOL
    if want_xent:
        assert deriv_err(x, r["xent"]) < DERIV_RTOL
    return res, d, x, r


def test_cfg2_full_size_matches_oracle():
    """BASELINE.json cfg 2, the configuration bench.py's headline is quoted on: P=2928, N=10 000,
    A~300k, B=128 x T=50, leaky 0.1, l2 5e-5 -- every number against the fp64 oracle, plus the
    size-independent properties the reference's own test asserts (chain-supervision-test.hpp:267-285)."""
    den, oden, g = _den("chime5")
    c = synth.CONFIGS["chime5"]
    res, d, x, r = _compare(den, oden, g, c["B"], c["T"], c["l2"], c["leaky"], seed=1235, want_xent=True)
    B, T = c["B"], c["T"]
    d64, x64 = d.astype(np.float64), x.astype(np.float64)
    y = synth.make_nnet_output(1236, B, T, den.num_pdfs, scale=2.0)
    mmi = d64 + c["l2"] * y                                    # derivative without the l2 part
    assert res[0] <= 0.0 and res[2] == B * T
    assert np.abs(mmi.sum(axis=1)).max() < 1e-3               # numerator and denominator posteriors sum to 1 per row
    assert np.abs(x64.sum(axis=1) - 1.0).max() < 1e-4
    den_post = x64 - mmi
    assert den_post.min() > -1e-5 and abs(den_post.sum() - B * T) < 1e-4 * B * T


def test_cfg2_graph_against_the_independent_autograd_formulation():
    """The CUDA path against tests/dense_ref.py's batched index_add + autograd formulation directly
    (the oracle is not involved): cfg 2's den graph, all T=50 frames, 8 sequences.  The CPU suite holds
    the oracle to the same formulation at this size (test_oracle_vs_independent_autograd_at_full_graph_size)."""
    import dense_ref
    den, oden, g = _den("chime5")
    c = synth.CONFIGS["chime5"]
    B, T = 8, c["T"]
    sup = synth.make_supervision(4101, den, B, T, num_paths=4, weight=1.0)
    y = synth.make_nnet_output(4102, B, T, den.num_pdfs, scale=2.0)
    objf, l2_term, w, ref = dense_ref.chain_objf_batched(den, oden.initial_probs(), sup, y, c["leaky"], c["l2"])
    s = io.Supervision.from_arrays(sup.weight, B, T, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                                   sup.arc_cost, sup.final_cost)
    res, d, _ = run_abi(g, s, y, c["l2"], c["leaky"])
    print("autograd parity: objf %.9g vs %.9g, deriv %.2e" % (res[0], objf, deriv_err(d, ref)))
    assert rel(res[0], objf) < OBJF_RTOL
    assert rel(res[1], l2_term) < OBJF_RTOL
    assert res[2] == pytest.approx(w)
    assert deriv_err(d, ref) < DERIV_RTOL


def test_cfg2_graph_weight_half_and_leaky_small():
    """Same graph, supervision.weight 0.5 and the reference's default leaky coefficient 1e-5
    (torchain/functions.py:129): the normalisers then carry the whole dynamic range."""
    den, oden, g = _den("chime5")
    _compare(den, oden, g, 128, 50, 0.0, 1e-5, seed=2001, weight=0.5)


def test_cfg2_graph_two_contexts_b256():
    """B = 256 (cfg 3 at 2 GPUs): two contexts of 128 sequences alternating inside one launch pair."""
    den, oden, g = _den("chime5")
    _compare(den, oden, g, 256, 50, 5e-5, 0.1, seed=2101)


def test_cfg2_graph_ragged_two_contexts_b200():
    """B = 200: the second context is partly padding."""
    den, oden, g = _den("chime5")
    _compare(den, oden, g, 200, 20, 5e-5, 0.1, seed=2151)


@pytest.mark.parametrize("B", [64, 32, 8])
def test_cfg2_graph_small_batches(B):
    """B = 64 / 32 / 8 (cfg 3 at 8 GPUs, cfg 5's small end): lane groups of 16 / 8 lanes (SUBS = 2 / 4)."""
    den, oden, g = _den("chime5")
    _compare(den, oden, g, B, 50, 5e-5, 0.1, seed=2200 + B)


def test_cfg2_graph_multi_pass_b512():
    """B = 512 on one GPU: two consecutive launch pairs of 256 (cfg 3 unsharded)."""
    den, oden, g = _den("chime5")
    _compare(den, oden, g, 512, 12, 5e-5, 0.1, seed=2301)


def test_cfg2_host_buffer_entry_full_size():
    """The e2e entry (host buffers, chunked copies overlapped with the recurrences) at cfg 2."""
    import torch

    def entry(g, s, y, l2, leaky, xent=0.0, want_xent=False):
        lib = _lib.load()
        rows, cols = y.shape
        yh = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).pin_memory()
        dh = torch.full((rows, cols), float("nan")).pin_memory()
        res = np.zeros(3, dtype=np.float32)
        for _ in range(2):   # the second call runs on the calibrated partition
            ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, yh.data_ptr(), rows, cols, cols,
                                                            res.ctypes.data_as(ctypes.c_void_p), dh.data_ptr(), cols,
                                                            None, 0, l2, leaky, 0.0, 0)
            assert ok == 1, _lib.last_error()
        return res, dh.numpy().copy(), None

    den, oden, g = _den("chime5")
    _compare(den, oden, g, 128, 50, 5e-5, 0.1, seed=2401, entry=entry)


def test_cfg4_full_size_matches_oracle():
    """BASELINE.json cfg 4 (large tree): P=6000, N=20 000, A~600k, B=256 x T=50."""
    den, oden, g = _den("large_tree")
    c = synth.CONFIGS["large_tree"]
    _compare(den, oden, g, c["B"], c["T"], c["l2"], c["leaky"], seed=3001)
    _den_cache.pop("large_tree")   # 5 GB of alpha store in the handle's workspace


def test_graph_whose_arc_slices_do_not_fit_shared_memory():
    """N=20 000 states x ~100 arcs: 2 M arcs = 13.6 k arcs (218 KB) per CTA, more than a CTA's shared
    memory, so both kernels read their arc slices from global memory (the unstaged instantiation) because
    they have to, not because an environment variable says so."""
    den = synth.make_den_fst(77, 20_000, 3000, mean_deg=100.0)
    assert den.num_arcs / 148 * 16 > 200 * 1024
    oden = ol.OracleDen(den)
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    _compare(den, oden, g, 128, 6, 5e-5, 0.1, seed=3101)


def test_cfg2_den_graph_read_from_an_openfst_file(tmp_path):
    """io.DenominatorGraph(path, n_pdf) -- the reference's constructor (torchain/io.py:51-54 ->
    src/my_lib_example.cpp:129-134: ReadFstKaldi + DenominatorGraph) -- on a cfg-2-sized binary den.fst
    (10 000 states, 300 k arcs, 4.9 MB): same graph, same initial probabilities, same loss as the handle built
    from arrays, and the loss matches the oracle."""
    import time
    den, oden, g_arrays = _den("chime5")
    path = str(tmp_path / "den.fst")
    synth.write_openfst_binary(path, den)
    t0 = time.perf_counter()
    g_file = io.DenominatorGraph(path, den.num_pdfs)
    print("read + prepared den.fst (%d bytes) in %.0f ms" % (os.path.getsize(path), 1e3 * (time.perf_counter() - t0)))
    assert (g_file.n_state, g_file.n_arc, g_file.n_pdf) == (den.num_states, den.num_arcs, den.num_pdfs)
    assert np.array_equal(g_file.initial_probs, g_arrays.initial_probs)
    res_f, d_f, _, r = _compare(den, oden, g_file, 128, 50, 5e-5, 0.1, seed=4001)
    sup = synth.make_supervision(4001, den, 128, 50, num_paths=4, weight=1.0)
    y = synth.make_nnet_output(4002, 128, 50, den.num_pdfs, scale=2.0)
    s = io.Supervision.from_arrays(sup.weight, 128, 50, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                                   sup.arc_cost, sup.final_cost)
    res_a, d_a, _ = run_abi(g_arrays, s, y, 5e-5, 0.1)
    assert rel(res_f[0], res_a[0]) < 1e-6 and deriv_err(d_f, d_a.astype(np.float64)) < 1e-5


def test_minibatch_from_chain_egs_bytes():
    """Supervisions pulled out of a binary egs ark (token scan, as /root/reference/example/chime5/ark2scp.py:24-40),
    merged into a minibatch (Kaldi's MergeChainExamples inside the reference's reader,
    src/my_lib_example_rand.cpp:160) and run through chain_loss: equal to the oracle on the appended FST."""
    import torch
    from test_abi import chain_egs_bytes
    from torchain.functions import chain_loss
    den = synth.make_den_fst(5, 800, 120, mean_deg=6.0)
    T = 30
    Bs = [1, 2, 1, 1, 3, 1, 2, 1]
    sups = [synth.make_supervision(600 + k, den, b, T, num_paths=3, weight=1.0) for k, b in enumerate(Bs)]
    ys = [synth.make_nnet_output(700 + k, b, T, den.num_pdfs, scale=1.5) for k, b in enumerate(Bs)]
    ark = chain_egs_bytes(sups, np.random.default_rng(9))
    minibatch = io.Supervision.merge(io.Supervision.from_chain_egs(ark))
    assert minibatch.shape == (sum(Bs), T, den.num_pdfs)
    y_all = synth.interleave_rows(ys, Bs)
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(synth.append_supervisions(sups)), y_all, l2=5e-5, leaky=0.1,
                 mode="fp64", threads=_cores())
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    inp = torch.from_numpy(y_all).cuda().requires_grad_(True)
    loss, results = chain_loss(inp, g, minibatch, l2_regularize=5e-5, leaky_hmm_coefficient=0.1)
    loss.backward()
    assert rel(float(results.data[0]), r["objf"]) < OBJF_RTOL
    assert deriv_err(inp.grad.cpu().numpy(), -r["deriv"]) < DERIV_RTOL


# ---- run-based ("tied") kernels: graphs whose pdfs belong to the state being entered -----------------------
@pytest.mark.parametrize("B,T", [(128, 50), (65, 20), (130, 12), (300, 8)])
def test_tied_cfg2_graph_matches_oracle(B, T):
    """config 2's shape with tied pdfs (10 000 states, 300 k arcs, ~20 k runs): den_tied_forward / backward kernels.
    B = 130 / 300 mix run-based passes of 128 with a per-arc pass for the remainder (2 / 44 sequences)."""
    den, oden, g = _den("chime5_tied")
    runs, _ = synth.in_pdf_runs(den)
    assert g.tied and g.n_run == runs and runs * 3 <= den.num_arcs
    assert not _den("chime5")[2].tied            # a pdf per arc: as many runs as arcs, per-arc kernels
    _compare(den, oden, g, B, T, 5e-5, 0.1, seed=5000 + B, want_xent=(B == 128))


def test_tied_and_per_arc_kernels_agree(monkeypatch):
    """The same tied graph through both kernel families (TORCHAIN_TIED=0 forces the per-arc kernels)."""
    den = synth.make_den_fst_tied(9, 3000, 400, mean_deg=20.0)
    oden = ol.OracleDen(den)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("TORCHAIN_TIED", mode)       # read once, when the graph handle is created
        g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                            den.final_cost, den.num_pdfs)
        out[mode] = _compare(den, oden, g, 100, 25, 5e-5, 1e-5, seed=5101, weight=0.5)
    assert deriv_err(out["1"][1], out["0"][1].astype(np.float64)) < 1e-5


def test_tied_kernels_on_degenerate_graph(monkeypatch):
    """States without in-arcs (alpha = 0 through the scratch run) and without out-arcs (beta' = 0), forced onto the
    run-based kernels."""
    monkeypatch.setenv("TORCHAIN_TIED", "1")
    src = np.array([0, 0, 1, 1, 2, 2, 3, 0], dtype=np.int32)
    dst = np.array([0, 1, 2, 0, 0, 4, 1, 2], dtype=np.int32)
    ilabel = np.array([1, 2, 3, 1, 1, 4, 2, 3], dtype=np.int32)    # tied: label = f(destination)
    prob = np.array([0.3, 0.3, 0.5, 0.4, 0.6, 0.3, 0.9, 0.3])
    den = synth.DenFst(5, 4, 0, src, dst, ilabel, (-np.log(prob)).astype(np.float32),
                       np.array([np.inf, 2.3, 2.3, 2.3, 0.1], dtype=np.float32))
    oden = ol.OracleDen(den)
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    B, T = 70, 6
    labels = [2, 3, 1]
    arc_begin, arc_dst, arc_label, arc_cost = [0], [], [], []
    for i in range(B * T):
        arc_dst.append(i + 1)
        arc_label.append(labels[(i % T) % 3])
        arc_cost.append(0.25)
        arc_begin.append(i + 1)
    arc_begin.append(B * T)
    final = np.full(B * T + 1, np.inf, dtype=np.float32)
    final[-1] = 0.0
    sup = synth.SupFst(1.0, B, T, 4, np.array(arc_begin, dtype=np.int32), np.array(arc_dst, dtype=np.int32),
                       np.array(arc_label, dtype=np.int32), np.array(arc_cost, dtype=np.float32), final)
    y = synth.make_nnet_output(5, B, T, 4)
    r = ol.chain(oden, ol.OracleSup(sup), y, l2=0.0, leaky=0.05, mode="fp64")
    s = io.Supervision.from_arrays(sup.weight, B, T, 4, sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost,
                                   sup.final_cost)
    res, d, _ = run_abi(g, s, y, 0.0, 0.05)
    assert rel(res[0], r["objf"]) < OBJF_RTOL
    assert deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_tied_host_buffer_entry():
    import torch
    den, oden, g = _den("chime5_tied")

    def entry(g, s, y, l2, leaky, xent=0.0, want_xent=False):
        lib = _lib.load()
        rows, cols = y.shape
        yh = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).pin_memory()
        dh = torch.full((rows, cols), float("nan")).pin_memory()
        res = np.zeros(3, dtype=np.float32)
        ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, yh.data_ptr(), rows, cols, cols,
                                                        res.ctypes.data_as(ctypes.c_void_p), dh.data_ptr(), cols,
                                                        None, 0, l2, leaky, 0.0, 0)
        assert ok == 1, _lib.last_error()
        return res, dh.numpy().copy(), None

    _compare(den, oden, g, 128, 30, 5e-5, 0.1, seed=5201, entry=entry)


def test_tied_host_buffer_entry_two_passes():
    """B = 200 on the tied graph through the host-buffer entry: two run-based passes (128 + 72 sequences), so the
    backward launches publish no progress and the derivative leaves after the compute stream has drained."""
    import torch
    den, oden, g = _den("chime5_tied")

    def entry(g, s, y, l2, leaky, xent=0.0, want_xent=False):
        lib = _lib.load()
        rows, cols = y.shape
        yh = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).pin_memory()
        dh = torch.full((rows, cols), float("nan")).pin_memory()
        res = np.zeros(3, dtype=np.float32)
        ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, yh.data_ptr(), rows, cols, cols,
                                                        res.ctypes.data_as(ctypes.c_void_p), dh.data_ptr(), cols,
                                                        None, 0, l2, leaky, 0.0, 0)
        assert ok == 1, _lib.last_error()
        return res, dh.numpy().copy(), None

    _compare(den, oden, g, 200, 10, 5e-5, 0.1, seed=5301, entry=entry)
