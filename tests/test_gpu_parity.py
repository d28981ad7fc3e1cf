"""GPU suite: the sm_100a path, called through the C ABI (ctypes) and through the
Python autograd op, against the oracle / golden fixtures on identical inputs."""
import ctypes

import numpy as np
import pytest

import oracle_lib as ol
from common import DERIV_RTOL, GOLDEN, OBJF_RTOL, deriv_err, load_golden, rel
from torchain_b200 import _lib, io, synth

pytestmark = pytest.mark.gpu


def handles(den, sup):
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    s = io.Supervision.from_arrays(sup.weight, sup.num_sequences, sup.frames_per_sequence, sup.label_dim,
                                   sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)
    return g, s


def run_abi(g, s, y_np, l2, leaky, xent=0.0, want_xent=False, stride_pad=0):
    """my_lib_ComputeChainObjfAndDeriv with raw device pointers."""
    import torch
    lib = _lib.load()
    rows, cols = y_np.shape
    y = torch.zeros(rows, cols + stride_pad, device="cuda")
    y[:, :cols] = torch.from_numpy(np.ascontiguousarray(y_np, dtype=np.float32)).cuda()
    deriv = torch.full((rows, cols + stride_pad), float("nan"), device="cuda")   # callee must not need zeroing
    xd = torch.full((rows, cols), float("nan"), device="cuda") if want_xent else None
    res = np.zeros(3, dtype=np.float32)
    ok = lib.my_lib_ComputeChainObjfAndDeriv(
        g.ptr, s.ptr, y.data_ptr(), rows, cols, y.stride(0), res.ctypes.data_as(ctypes.c_void_p),
        deriv.data_ptr(), deriv.stride(0), xd.data_ptr() if want_xent else None, cols,
        l2, leaky, xent, torch.cuda.current_device(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert ok == 1, _lib.last_error()
    torch.cuda.synchronize()
    return res, deriv[:, :cols].cpu().numpy(), (xd.cpu().numpy() if want_xent else None)


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p[-10:-4])
def test_gpu_matches_golden(path):
    c = load_golden(path)
    g, s = handles(c["den"], c["sup"])
    res, d, _ = run_abi(g, s, c["y"], c["l2"], c["leaky"])
    assert rel(res[0], c["objf"]) < OBJF_RTOL
    assert abs(res[1] - c["l2_term"]) <= OBJF_RTOL * max(abs(c["l2_term"]), 1e-30)
    assert res[2] == pytest.approx(c["weight"])
    assert deriv_err(d, c["deriv"]) < DERIV_RTOL


@pytest.mark.parametrize("seed", range(50))
def test_gpu_matches_oracle_small(seed):
    """cfg 1 of BASELINE.json: the reference's own CPU-runnable random case ranges."""
    c = synth.small_case(seed)
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    r = ol.chain(oden, osup, c["y"], l2=c["l2"], leaky=c["leaky"], mode="fp64")
    g, s = handles(c["den"], c["sup"])
    res, d, _ = run_abi(g, s, c["y"], c["l2"], c["leaky"], stride_pad=seed % 3)
    assert rel(res[0], r["objf"]) < OBJF_RTOL
    assert abs(res[1] - r["l2_term"]) <= OBJF_RTOL * max(abs(r["l2_term"]), 1e-30)
    assert res[2] == pytest.approx(r["weight"])
    assert deriv_err(d, r["deriv"]) < DERIV_RTOL
    # property assertions of the reference's test (chain-supervision-test.hpp:267-285)
    if c["l2"] == 0.0:
        assert np.linalg.norm(d.astype(np.float64).sum(axis=1)) < 0.1
        assert res[0] <= 0.0


@pytest.mark.parametrize("B,T,N,P", [(33, 9, 300, 70), (64, 11, 500, 100), (100, 7, 400, 64),
                                     (128, 6, 700, 130), (200, 5, 350, 90), (300, 4, 250, 50)])
def test_gpu_matches_oracle_batch_shapes(B, T, N, P):
    """Every vector width / chunking / multi-pass path of the kernel (padded batches)."""
    den = synth.make_den_fst(B, N, P, mean_deg=6.0)
    sup = synth.make_supervision(B + 1, den, B, T, num_paths=3, weight=1.0)
    y = synth.make_nnet_output(B + 2, B, T, P, scale=1.5)
    oden, osup = ol.OracleDen(den), ol.OracleSup(sup)
    r = ol.chain(oden, osup, y, l2=5e-5, leaky=0.1, mode="fp64", threads=4)
    g, s = handles(den, sup)
    res, d, _ = run_abi(g, s, y, 5e-5, 0.1)
    assert rel(res[0], r["objf"]) < OBJF_RTOL
    assert rel(res[1], r["l2_term"]) < OBJF_RTOL
    assert deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_gpu_xent_output_is_weighted_numerator_posterior():
    c = synth.small_case(7)
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    r = ol.chain(oden, osup, c["y"], l2=c["l2"], leaky=c["leaky"], mode="fp64", want_xent=True)
    g, s = handles(c["den"], c["sup"])
    res, d, x = run_abi(g, s, c["y"], c["l2"], c["leaky"], xent=0.1, want_xent=True)
    assert deriv_err(x, r["xent"]) < DERIV_RTOL
    assert deriv_err(d, r["deriv"]) < DERIV_RTOL
    # every (t, s) row of the numerator posterior sums to supervision.weight
    assert np.abs(x.astype(np.float64).sum(axis=1) - c["sup"].weight).max() < 1e-4


def test_gpu_unsplit_supervision_fallback():
    """A merged FST that does not factorise per sequence runs as one chain and still matches."""
    den = synth.make_den_fst(3, 50, 12, mean_deg=3.0)
    sup = synth.make_supervision(4, den, 3, 6, num_paths=3)
    times = np.full(sup.num_states, -1)
    times[0] = 0
    for st in range(sup.num_states):
        for a in range(sup.arc_begin[st], sup.arc_begin[st + 1]):
            times[sup.arc_dst[a]] = times[st] + 1
    boundary = np.where(times == 6)[0]
    if len(boundary) < 2:
        pytest.skip("needs two boundary states")
    sup.arc_cost = sup.arc_cost.copy()
    sup.arc_cost[sup.arc_begin[boundary[1]]] += 0.5
    y = synth.make_nnet_output(5, 3, 6, 12)
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=0.0, leaky=0.1, mode="fp64")
    g, s = handles(den, sup)
    assert s.num_chains == 1
    res, d, _ = run_abi(g, s, y, 0.0, 0.1)
    assert rel(res[0], r["objf"]) < OBJF_RTOL
    assert deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_gpu_guard_on_nan():
    c = synth.small_case(1)
    g, s = handles(c["den"], c["sup"])
    y = c["y"].copy()
    y[0, 0] = np.nan
    res, d, _ = run_abi(g, s, y, 0.0, 0.1)
    assert res[0] == pytest.approx(-10.0 * res[2])
    assert np.all(d == 0.0)


def test_gpu_error_paths():
    import torch
    lib = _lib.load()
    c = synth.small_case(2)
    g, s = handles(c["den"], c["sup"])
    y = torch.from_numpy(c["y"]).cuda()
    d = torch.empty_like(y)
    res = np.zeros(3, dtype=np.float32)
    args = lambda rows, leaky: (g.ptr, s.ptr, y.data_ptr(), rows, y.shape[1], y.stride(0),
                                res.ctypes.data_as(ctypes.c_void_p), d.data_ptr(), d.stride(0), None, 0,
                                0.0, leaky, 0.0, 0, None)
    assert lib.my_lib_ComputeChainObjfAndDeriv(*args(y.shape[0] - 1, 0.1)) == 0
    assert "rows" in _lib.last_error()
    assert lib.my_lib_ComputeChainObjfAndDeriv(*args(y.shape[0], 0.0)) == 0
    assert "leaky" in _lib.last_error()


def test_python_chain_loss_autograd():
    """torchain.functions.chain_loss semantics (reference functions.py:62-138)."""
    import torch
    from torchain_b200.functions import chain_loss
    c = synth.small_case(9)
    B, T, P = c["B"], c["T"], c["P"]
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    r = ol.chain(oden, osup, c["y"], l2=0.01, leaky=0.1, mode="fp64", want_xent=True)
    g, s = handles(c["den"], c["sup"])
    # (B, P, T) input as the recipes pass it
    y3 = torch.from_numpy(c["y"]).view(T, B, P).permute(1, 2, 0).contiguous().cuda().requires_grad_(True)
    x3 = y3.detach().clone().requires_grad_(True)
    loss, results = chain_loss(y3, g, s, l2_regularize=0.01, leaky_hmm_coefficient=0.1,
                               xent_regularize=0.1, xent_input=x3, kaldi_way=True)
    assert loss.is_cuda and loss.shape == (1,)
    assert float(loss) == pytest.approx(-r["objf"] / r["weight"], rel=1e-5)
    assert not results.data.is_cuda and results.data.dtype == torch.float32
    assert float(results.data[0]) == pytest.approx(r["objf"], rel=1e-5)
    assert float(results.data[1]) == pytest.approx(r["l2_term"], rel=1e-5)
    (loss * 123.0).backward()   # grad_output is ignored by the reference's backward
    gy = y3.grad.permute(2, 0, 1).reshape(T * B, P).cpu().numpy()
    gx = x3.grad.permute(2, 0, 1).reshape(T * B, P).cpu().numpy()
    assert deriv_err(gy, -r["deriv"]) < DERIV_RTOL
    assert deriv_err(gx, -0.1 * r["xent"]) < DERIV_RTOL


def test_host_buffer_entry():
    lib = _lib.load()
    c = load_golden(GOLDEN[2])
    g, s = handles(c["den"], c["sup"])
    y = np.ascontiguousarray(c["y"], dtype=np.float32)
    d = np.full_like(y, np.nan)
    res = np.zeros(3, dtype=np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, p(y), y.shape[0], y.shape[1], y.shape[1], p(res),
                                                    p(d), y.shape[1], None, 0, c["l2"], c["leaky"], 0.0, 0)
    assert ok == 1, _lib.last_error()
    assert rel(res[0], c["objf"]) < OBJF_RTOL
    assert deriv_err(d, c["deriv"]) < DERIV_RTOL


def _host_entry(g, s, y, l2, leaky, pinned=True):
    """torchain_ComputeChainObjfAndDeriv_host on (pinned) host buffers; returns results, derivative."""
    import torch
    lib = _lib.load()
    rows, cols = y.shape
    yh = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
    dh = torch.full((rows, cols), float("nan"))
    if pinned:
        yh, dh = yh.pin_memory(), dh.pin_memory()
    res = np.zeros(3, dtype=np.float32)
    ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, yh.data_ptr(), rows, cols, cols,
                                                    res.ctypes.data_as(ctypes.c_void_p), dh.data_ptr(), cols,
                                                    None, 0, l2, leaky, 0.0, 0)
    assert ok == 1, _lib.last_error()
    return res, dh.numpy().copy()


@pytest.mark.parametrize("T", [1, 2, 17, 50, 77])
def test_host_buffer_entry_chunk_schedules(T):
    """The host-buffer form moves the minibatch in chunks of whole frames of DIFFERENT sizes (single
    frames at both ends, torchain_host_chunk_bounds); every schedule must deliver what the device form
    computes from the same input."""
    den = synth.make_den_fst(5, 1500, 200, mean_deg=8.0)
    B = 48
    sup = synth.make_supervision(6, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(7, B, T, den.num_pdfs, scale=1.5)
    g, s = handles(den, sup)
    r_dev, d_dev, _ = run_abi(g, s, y, 5e-5, 0.1)
    r_host, d_host = _host_entry(g, s, y, 5e-5, 0.1)
    assert rel(r_host[0], r_dev[0]) < OBJF_RTOL and rel(r_host[1], r_dev[1]) < OBJF_RTOL and r_host[2] == r_dev[2]
    assert np.isfinite(d_host).all()
    assert deriv_err(d_host, d_dev.astype(np.float64)) < 1e-5


def test_host_buffer_entry_guard_rewrites_early_chunks():
    """When the numeric guard trips, the final pass rewrites the derivative (l2 part only) AFTER the
    host-buffer form has already copied the chunks of the late frames out during the backward launch:
    those chunks have to be delivered again (Kaldi: objf = -10 * weight, deriv = -w * l2 * nnet_output)."""
    den = synth.make_den_fst(8, 4000, 300, mean_deg=12.0)
    B, T = 128, 60  # a backward launch long enough for chunks to leave early
    sup = synth.make_supervision(9, den, B, T, num_paths=2, weight=0.5)
    y = synth.make_nnet_output(10, B, T, den.num_pdfs, scale=1.0)
    y[3, 7] = np.nan
    g, s = handles(den, sup)
    l2 = 0.01
    for _ in range(2):
        res, d = _host_entry(g, s, y, l2, 0.1)
        assert res[0] == pytest.approx(-10.0 * res[2]) and res[2] == pytest.approx(0.5 * B * T)
        want = -0.5 * l2 * y
        fin = np.isfinite(y)
        assert np.isnan(d[~fin]).all()
        np.testing.assert_allclose(d[fin], want[fin], rtol=1e-6, atol=0)


def test_gpu_single_frame_and_single_sequence():
    """T = 1 (the reference's own test skips it, chain-supervision-test.hpp:243) and B = 1."""
    for (B, T) in [(1, 1), (3, 1), (1, 5)]:
        den = synth.make_den_fst(50 + B + T, 30, 9, mean_deg=2.0)
        sup = synth.make_supervision(60, den, B, T, num_paths=2)
        y = synth.make_nnet_output(61, B, T, 9)
        r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=0.01, leaky=0.1, mode="fp64")
        g, s = handles(den, sup)
        res, d, _ = run_abi(g, s, y, 0.01, 0.1)
        assert rel(res[0], r["objf"]) < OBJF_RTOL
        assert deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_gpu_states_without_in_or_out_arcs():
    """Hand-built graph: state 3 has no in-arcs, state 4 has no out-arcs (only a final weight)."""
    src = np.array([0, 0, 1, 1, 2, 2, 3, 0], dtype=np.int32)
    dst = np.array([0, 1, 2, 0, 0, 4, 1, 2], dtype=np.int32)
    ilabel = np.array([1, 2, 3, 1, 2, 4, 3, 4], dtype=np.int32)
    prob = np.array([0.3, 0.3, 0.5, 0.4, 0.6, 0.3, 0.9, 0.3])
    den = synth.DenFst(5, 4, 0, src, dst, ilabel, (-np.log(prob)).astype(np.float32),
                       np.array([np.inf, 2.3, 2.3, 2.3, 0.1], dtype=np.float32))
    B, T = 4, 6
    # supervision: every sequence follows 0 -> 1 -> 2 -> 0 -> ... (labels 2, 3, 2) with unit costs
    labels = [2, 3, 2]
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
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=0.0, leaky=0.05, mode="fp64")
    g, s = handles(den, sup)
    assert s.num_chains == B
    res, d, _ = run_abi(g, s, y, 0.0, 0.05)
    assert rel(res[0], r["objf"]) < OBJF_RTOL
    assert deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_python_chain_loss_not_kaldi_way():
    """kaldi_way=False (reference functions.py:96-103): the loss runs a second time on xent_input and
    overwrites BOTH results and the LF-MMI gradient buffer."""
    import torch
    from torchain_b200.functions import chain_loss
    c = synth.small_case(11)
    B, T, P = c["B"], c["T"], c["P"]
    g, s = handles(c["den"], c["sup"])
    y = torch.from_numpy(c["y"]).cuda().requires_grad_(True)
    x_np = synth.make_nnet_output(99, B, T, P)
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    r = ol.chain(ol.OracleDen(c["den"]), ol.OracleSup(c["sup"]), x_np, l2=0.0, leaky=0.1, mode="fp64", want_xent=True)
    loss, results = chain_loss(y, g, s, leaky_hmm_coefficient=0.1, xent_regularize=0.25, xent_input=x, kaldi_way=False)
    loss.backward()
    assert float(results.data[0]) == pytest.approx(r["objf"], rel=1e-5)        # results of the SECOND call
    assert deriv_err(y.grad.cpu().numpy(), -r["deriv"]) < DERIV_RTOL             # mmi grad overwritten too
    assert deriv_err(x.grad.cpu().numpy(), -0.25 * r["xent"]) < DERIV_RTOL


def _hub_graph(seed, N, P, hubs):
    """Random graph plus `hubs` states that every state has an arc to and from (heavy-tailed degrees)."""
    rng = np.random.default_rng(seed)
    den = synth.make_den_fst(seed, N, P, mean_deg=3.0)
    src, dst, lab, cost = [den.src], [den.dst], [den.ilabel], [den.cost]
    for h in range(hubs):
        others = np.arange(N, dtype=np.int32)
        src += [others, np.full(N, h, dtype=np.int32)]
        dst += [np.full(N, h, dtype=np.int32), others]
        lab += [rng.integers(1, P + 1, size=N).astype(np.int32)] * 2
        cost += [np.full(N, 3.0, dtype=np.float32), np.full(N, np.log(4.0 * N), dtype=np.float32)]
    src, dst, lab, cost = map(np.concatenate, (src, dst, lab, cost))
    o = np.argsort(src, kind="stable")
    return synth.DenFst(N, P, 0, src[o], dst[o], lab[o], cost[o], den.final_cost)


@pytest.mark.parametrize("cap", [None, "7"])
def test_gpu_split_high_degree_states(cap, monkeypatch):
    """States with more arcs than the split cap are spread over several warp tasks that accumulate
    with REDs (default cap 192; a tiny cap also splits ordinary states)."""
    if cap is not None:
        monkeypatch.setenv("TORCHAIN_SPLIT_CAP", cap)
    den = _hub_graph(5, 300, 40, hubs=2)
    B, T = 37, 7
    sup = synth.make_supervision(6, den, B, T, num_paths=3)
    y = synth.make_nnet_output(7, B, T, 40, scale=1.5)
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=0.01, leaky=0.1, mode="fp64", threads=4)
    g, s = handles(den, sup)
    for _ in range(2):   # second call reuses (un-zeroed) workspace buffers
        res, d, _ = run_abi(g, s, y, 0.01, 0.1)
        assert rel(res[0], r["objf"]) < OBJF_RTOL
        assert deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_gpu_ex_entry_async_results_and_objective_only():
    """torchain_ComputeChainObjfAndDeriv_ex: results to a device triple without a host sync, and
    nnet_output_deriv = NULL (objective only, as Kaldi allows)."""
    import torch
    lib = _lib.load()
    c = load_golden(GOLDEN[1])
    g, s = handles(c["den"], c["sup"])
    y = torch.from_numpy(np.ascontiguousarray(c["y"], dtype=np.float32)).cuda()
    res_dev = torch.zeros(3, device="cuda")
    ok = lib.torchain_ComputeChainObjfAndDeriv_ex(
        g.ptr, s.ptr, y.data_ptr(), y.shape[0], y.shape[1], y.stride(0), None, res_dev.data_ptr(),
        None, 0, None, 0, c["l2"], c["leaky"], 0.0, 1.0, 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert ok == 1, _lib.last_error()
    r = res_dev.cpu().numpy()
    assert rel(r[0], c["objf"]) < OBJF_RTOL and r[2] == pytest.approx(c["weight"])
    assert abs(r[1] - c["l2_term"]) <= OBJF_RTOL * max(abs(c["l2_term"]), 1e-30)


def test_gpu_deriv_scale_and_repeatability():
    """deriv_scale = -1 yields the negated derivative (what backward() returns); two calls agree to
    rounding (the occupancy REDs make the summation order vary, nothing else)."""
    import torch
    lib = _lib.load()
    c = synth.small_case(13)
    g, s = handles(c["den"], c["sup"])
    y = torch.from_numpy(c["y"]).cuda()
    outs = []
    for scale in (1.0, -1.0, -1.0):
        d = torch.empty_like(y)
        res = np.zeros(3, dtype=np.float32)
        ok = lib.torchain_ComputeChainObjfAndDeriv_ex(
            g.ptr, s.ptr, y.data_ptr(), y.shape[0], y.shape[1], y.stride(0), res.ctypes.data_as(ctypes.c_void_p), None,
            d.data_ptr(), d.stride(0), None, 0, 0.01, 0.1, 0.0, scale, 0, None)
        assert ok == 1, _lib.last_error()
        outs.append(d.cpu().numpy().astype(np.float64))
    assert np.abs(outs[0] + outs[1]).max() <= 1e-5 * np.abs(outs[0]).max()
    assert np.abs(outs[1] - outs[2]).max() <= 1e-5 * np.abs(outs[0]).max()


def test_gpu_supervision_from_kaldi_bytes_matches_arrays():
    from test_abi import kaldi_supervision_bytes
    c = load_golden(GOLDEN[2])
    g, _ = handles(c["den"], c["sup"])
    s, _ = io.Supervision.from_kaldi_bytes(kaldi_supervision_bytes(c["sup"]))
    res, d, _ = run_abi(g, s, c["y"], c["l2"], c["leaky"])
    assert rel(res[0], c["objf"]) < OBJF_RTOL
    assert deriv_err(d, c["deriv"]) < DERIV_RTOL


def test_gpu_unstaged_arcs_path(monkeypatch):
    """Graphs whose per-CTA arc slice does not fit in shared memory read arcs from global memory."""
    monkeypatch.setenv("TORCHAIN_FORCE_UNSTAGED", "1")
    for seed, B in ((21, 5), (22, 130)):
        den = synth.make_den_fst(seed, 200, 30, mean_deg=4.0)
        sup = synth.make_supervision(seed + 1, den, B, 6, num_paths=3)
        y = synth.make_nnet_output(seed + 2, B, 6, 30)
        r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=0.0, leaky=0.1, mode="fp64", threads=4)
        g, s = handles(den, sup)
        res, d, _ = run_abi(g, s, y, 0.0, 0.1)
        assert rel(res[0], r["objf"]) < OBJF_RTOL
        assert deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_gpu_degenerate_graphs_and_leaky_extremes():
    """One-state / one-pdf graph, and leaky coefficients near both ends of (0, 1)."""
    den1 = synth.DenFst(1, 1, 0, np.array([0], dtype=np.int32), np.array([0], dtype=np.int32),
                        np.array([1], dtype=np.int32), np.array([0.1], dtype=np.float32),
                        np.array([2.0], dtype=np.float32))
    B, T = 3, 4
    S = B * T + 1
    sup1 = synth.SupFst(1.0, B, T, 1, np.arange(S + 1, dtype=np.int32).clip(max=S - 1),
                        np.arange(1, S, dtype=np.int32), np.ones(S - 1, dtype=np.int32),
                        np.full(S - 1, 0.3, dtype=np.float32),
                        np.concatenate([np.full(S - 1, np.inf), [0.0]]).astype(np.float32))
    y1 = synth.make_nnet_output(1, B, T, 1)
    cases = [(den1, sup1, y1, 0.1)]
    den = synth.make_den_fst(31, 70, 13, mean_deg=3.0)
    sup = synth.make_supervision(32, den, 4, 9, num_paths=3)
    y = synth.make_nnet_output(33, 4, 9, 13, scale=2.0)
    cases += [(den, sup, y, 1e-6), (den, sup, y, 0.9)]
    for (dn, sp, yy, leaky) in cases:
        r = ol.chain(ol.OracleDen(dn), ol.OracleSup(sp), yy, l2=0.0, leaky=leaky, mode="fp64")
        g, s = handles(dn, sp)
        res, d, _ = run_abi(g, s, yy, 0.0, leaky)
        assert rel(res[0], r["objf"]) < OBJF_RTOL
        # the one-pdf case has an exactly zero derivative (both posteriors are 1): absolute tolerance there
        assert np.abs(d.astype(np.float64) - r["deriv"]).max() < DERIV_RTOL * max(1.0, np.abs(r["deriv"]).max())


def test_gpu_two_devices_one_process():
    """The reference's parallel_train.py drives several GPUs from one process; handles keep
    per-device copies of the graph and per-(device, stream) workspaces."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    lib = _lib.load()
    c = load_golden(GOLDEN[2])
    g, s = handles(c["den"], c["sup"])
    outs = []
    for dev in (0, 1, 0):
        with torch.cuda.device(dev):
            y = torch.from_numpy(np.ascontiguousarray(c["y"], dtype=np.float32)).cuda(dev)
            d = torch.empty_like(y)
            res = np.zeros(3, dtype=np.float32)
            ok = lib.my_lib_ComputeChainObjfAndDeriv(
                g.ptr, s.ptr, y.data_ptr(), y.shape[0], y.shape[1], y.stride(0), res.ctypes.data_as(ctypes.c_void_p),
                d.data_ptr(), d.stride(0), None, 0, c["l2"], c["leaky"], 0.0, dev,
                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            assert ok == 1, _lib.last_error()
            torch.cuda.synchronize(dev)
            outs.append((res.copy(), d.cpu().numpy()))
    for res, d in outs:
        assert rel(res[0], c["objf"]) < OBJF_RTOL
        assert deriv_err(d, c["deriv"]) < DERIV_RTOL


def test_gpu_two_threads_two_streams():
    """Re-entrancy per (handle, stream): two host threads drive the same handles on their own
    CUDA streams (the reference's parallel_apply used Python threads, parallel_train.py:67)."""
    import threading
    import torch
    lib = _lib.load()
    c = synth.make_config("chime5", B=32, T=8, with_y=False)
    g, s = handles(c["den"], c["sup"])
    y = torch.randn(8 * 32, c["P"], device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    ref_d = torch.empty_like(y)
    ref = np.zeros(3, dtype=np.float32)
    args = lambda d, res, st: (g.ptr, s.ptr, y.data_ptr(), y.shape[0], y.shape[1], y.stride(0),
                               res.ctypes.data_as(ctypes.c_void_p), d.data_ptr(), d.stride(0), None, 0,
                               0.0, 0.1, 0.0, 0, ctypes.c_void_p(st.cuda_stream))
    assert lib.my_lib_ComputeChainObjfAndDeriv(*args(ref_d, ref, torch.cuda.current_stream())) == 1
    torch.cuda.synchronize()
    results = {}

    def worker(i):
        st = torch.cuda.Stream()
        d = torch.empty_like(y)
        res = np.zeros(3, dtype=np.float32)
        st.wait_stream(torch.cuda.default_stream())
        for _ in range(5):
            ok = lib.my_lib_ComputeChainObjfAndDeriv(*args(d, res, st))
            assert ok == 1, _lib.last_error()
        st.synchronize()
        results[i] = (res.copy(), d.clone())

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
        assert not t.is_alive()
    for i in range(2):
        res, d = results[i]
        assert rel(res[0], ref[0]) < OBJF_RTOL
        assert float((d - ref_d).abs().max()) <= 1e-4 * float(ref_d.abs().max())


# ---------------------------------------------------------------- strided (to2d-free) entry
def run_strided(g, s, y_np, B, T, l2, leaky, layout, xent=0.0, want_xent=False, deriv_scale=1.0):
    """torchain_ComputeChainObjfAndDeriv_strided on a device tensor laid out as `layout`:
    "BPT" = the network's native (B, P, T); "BTP" = (B, T, P); "TBP" = the reference's 2-D form.
    Returns results and the derivative(s) brought back to the (T*B, P) order of the oracle."""
    import torch
    lib = _lib.load()
    P = y_np.shape[1]
    y_tbp = torch.from_numpy(np.ascontiguousarray(y_np, dtype=np.float32)).cuda().view(T, B, P)
    perm = {"BPT": (1, 2, 0), "BTP": (1, 0, 2), "TBP": (0, 1, 2)}[layout]
    y = y_tbp.permute(*perm).contiguous()
    names = layout

    def strides(x):
        return (ctypes.c_int64 * 3)(x.stride(names.index("T")), x.stride(names.index("B")), x.stride(names.index("P")))

    deriv = torch.full_like(y, float("nan"))
    xd = torch.full_like(y, float("nan")) if want_xent else None
    res = np.zeros(3, dtype=np.float32)
    ok = lib.torchain_ComputeChainObjfAndDeriv_strided(
        g.ptr, s.ptr, y.data_ptr(), T, B, P, strides(y), res.ctypes.data_as(ctypes.c_void_p), None,
        deriv.data_ptr(), strides(deriv), xd.data_ptr() if want_xent else None, strides(xd) if want_xent else None,
        l2, leaky, xent, deriv_scale, torch.cuda.current_device(),
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert ok == 1, _lib.last_error()
    torch.cuda.synchronize()
    inv = [perm.index(i) for i in range(3)]
    back = lambda x: x.permute(*inv).reshape(T * B, P).cpu().numpy()
    return res, back(deriv), (back(xd) if want_xent else None)


@pytest.mark.parametrize("layout", ["BPT", "BTP", "TBP"])
@pytest.mark.parametrize("seed", [0, 3, 5, 11, 18, 26])
def test_gpu_strided_layouts_match_oracle(seed, layout):
    """Row f4 of SURVEY.md section 8: the network output read in its own (B, P, T) layout and the
    gradient written there, no to2d copy; same numbers as the 2-D path."""
    c = synth.small_case(seed)
    sup = c["sup"]
    r = ol.chain(ol.OracleDen(c["den"]), ol.OracleSup(sup), c["y"], l2=c["l2"], leaky=c["leaky"], mode="fp64",
                 want_xent=True)
    g, s = handles(c["den"], sup)
    res, d, x = run_strided(g, s, c["y"], sup.num_sequences, sup.frames_per_sequence, c["l2"], c["leaky"], layout,
                            xent=0.1, want_xent=True)
    assert rel(res[0], r["objf"]) < OBJF_RTOL
    assert abs(res[1] - r["l2_term"]) <= OBJF_RTOL * max(abs(r["l2_term"]), 1e-30)
    assert res[2] == pytest.approx(r["weight"])
    assert deriv_err(d, r["deriv"]) < DERIV_RTOL
    assert deriv_err(x, r["xent"]) < DERIV_RTOL


@pytest.mark.parametrize("B,T,N,P", [(64, 11, 500, 100), (128, 9, 700, 130), (200, 5, 350, 90)])
def test_gpu_strided_bpt_batch_shapes(B, T, N, P):
    den = synth.make_den_fst(B, N, P, mean_deg=6.0)
    sup = synth.make_supervision(B + 1, den, B, T, num_paths=3, weight=0.5)
    y = synth.make_nnet_output(B + 2, B, T, P, scale=1.5)
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=5e-5, leaky=0.1, mode="fp64", threads=4)
    g, s = handles(den, sup)
    res, d, _ = run_strided(g, s, y, B, T, 5e-5, 0.1, "BPT", deriv_scale=-1.0)
    assert rel(res[0], r["objf"]) < OBJF_RTOL
    assert rel(res[1], r["l2_term"]) < OBJF_RTOL
    assert deriv_err(-d, r["deriv"]) < DERIV_RTOL


def test_gpu_strided_guard_on_nan():
    """Kaldi's numeric guard rewrites the whole derivative: it must follow the caller's layout too."""
    c = synth.small_case(1)
    sup = c["sup"]
    g, s = handles(c["den"], sup)
    y = c["y"].copy()
    y[0, 0] = np.nan
    res, d, x = run_strided(g, s, y, sup.num_sequences, sup.frames_per_sequence, 0.0, 0.1, "BPT", xent=0.1,
                            want_xent=True)
    assert res[0] == pytest.approx(-10.0 * res[2])
    assert np.all(d == 0.0) and np.all(x == 0.0)


def test_gpu_strided_error_paths():
    import torch
    lib = _lib.load()
    c = synth.small_case(2)
    sup = c["sup"]
    g, s = handles(c["den"], sup)
    B, T, P = sup.num_sequences, sup.frames_per_sequence, c["y"].shape[1]
    y = torch.zeros(B, P, T, device="cuda")
    d = torch.empty_like(y)
    res = np.zeros(3, dtype=np.float32)
    good = (ctypes.c_int64 * 3)(1, P * T, T)
    zero = (ctypes.c_int64 * 3)(1, 0, T)
    call = lambda nt, ys, ds: lib.torchain_ComputeChainObjfAndDeriv_strided(
        g.ptr, s.ptr, y.data_ptr(), nt, B, P, ys, res.ctypes.data_as(ctypes.c_void_p), None, d.data_ptr(), ds,
        None, None, 0.0, 0.1, 0.0, 1.0, 0, None)
    assert call(T + 1, good, good) == 0 and "frames" in _lib.last_error()
    assert call(T, None, good) == 0 and "stride" in _lib.last_error()
    assert call(T, good, zero) == 0 and "positive" in _lib.last_error()
    assert call(T, good, good) == 1, _lib.last_error()
    torch.cuda.synchronize()


def test_gpu_chain_loss_3d_input_needs_no_to2d():
    """chain_loss on the (B, P, T) network output equals the reference's to2d route
    (functions.py:118-138) in loss, results and the gradient reaching the 3-D tensor."""
    import torch
    from torchain_b200.functions import chain_loss, to2d
    c = synth.small_case(9)
    sup = c["sup"]
    B, T, P = sup.num_sequences, sup.frames_per_sequence, c["y"].shape[1]
    g, s = handles(c["den"], sup)
    x3 = torch.from_numpy(c["y"]).cuda().view(T, B, P).permute(1, 2, 0).contiguous()
    xe3 = (0.5 * x3 + 0.1).contiguous()

    def run(use_to2d, **kw):
        a = x3.clone().requires_grad_(True)
        b = xe3.clone().requires_grad_(True)
        loss, results = chain_loss(to2d(a) if use_to2d else a, g, s, l2_regularize=0.01, leaky_hmm_coefficient=0.1,
                                   xent_regularize=0.1, xent_input=to2d(b) if use_to2d else b, **kw)
        loss.backward()
        return float(loss), results.data.clone(), a.grad.cpu().numpy(), b.grad.cpu().numpy()

    for kw in ({"kaldi_way": True}, {"kaldi_way": False}):
        l2d, r2d, ga2, gb2 = run(True, **kw)
        l3d, r3d, ga3, gb3 = run(False, **kw)
        assert l3d == pytest.approx(l2d, rel=1e-6)
        assert torch.allclose(r3d, r2d, rtol=1e-6)
        assert ga3.shape == (B, P, T) and gb3.shape == (B, P, T)
        assert deriv_err(ga3, ga2.astype(np.float64)) < 1e-5
        assert deriv_err(gb3, gb2.astype(np.float64)) < 1e-5
    # a (B, T, P) tensor viewed as (B, P, T): arbitrary strides are read in place as well
    v = x3.permute(0, 2, 1).contiguous().permute(0, 2, 1).requires_grad_(True)
    assert not v.is_contiguous()
    loss, results = chain_loss(v, g, s, l2_regularize=0.01, leaky_hmm_coefficient=0.1)
    loss.backward()
    a = x3.clone().requires_grad_(True)
    loss2, _ = chain_loss(to2d(a), g, s, l2_regularize=0.01, leaky_hmm_coefficient=0.1)
    loss2.backward()
    assert float(loss) == pytest.approx(float(loss2), rel=1e-6)
    assert deriv_err(v.grad.cpu().numpy(), a.grad.cpu().numpy().astype(np.float64)) < 1e-5


@pytest.mark.parametrize("B,P,T", [(1, 1, 2), (3, 5, 7), (4, 33, 50), (2, 64, 256), (5, 70, 300), (128, 96, 50),
                                   (2, 31, 1599), (128, 2928, 50), (7, 48, 150), (3, 40, 13), (2, 36, 4), (9, 2928, 1)])
def test_gpu_relayout_kernels_equal_permute(B, P, T):
    """torchain_to2d / torchain_from2d are bit-exact copies of the reference's to2d
    (functions.py:118-125) and of the permute its backward applies to the gradient.  Shapes whose blocks and rows
    are 16-byte aligned (P*T and the last block's pdfs*T multiples of 4, row stride a multiple of 4, T <= 256) go
    through the TMA kernels (bulk copy + 3-D tensor copy), every other shape -- and the padded stride P+3 below --
    through the plain ones; both are covered, and a padded stride of P+4 sends the same data through TMA."""
    import torch
    from torchain_b200.functions import to2d
    x = torch.randn(B, P, T, device="cuda")
    want = x.permute(2, 0, 1).contiguous().view(-1, P)
    xg = x.clone().requires_grad_(True)
    rows = to2d(xg)
    assert rows.shape == (T * B, P) and torch.equal(rows, want)
    w = torch.randn_like(rows)
    (rows * w).sum().backward()
    assert torch.equal(xg.grad, w.view(T, B, P).permute(1, 2, 0))
    # padded row stride and scaling through the C ABI
    lib = _lib.load()
    pad = torch.full((T * B, P + 3), 7.0, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.torchain_to2d(x.data_ptr(), B, P, T, pad.data_ptr(), P + 3, 0, st) == 1, _lib.last_error()
    assert torch.equal(pad[:, :P], want) and bool((pad[:, P:] == 7.0).all())
    back = torch.empty_like(x)
    assert lib.torchain_from2d(pad.data_ptr(), P + 3, B, P, T, back.data_ptr(), -2.0, 0, st) == 1, _lib.last_error()
    assert torch.equal(back, -2.0 * x)
    assert lib.torchain_to2d(x.data_ptr(), B, P, T, pad.data_ptr(), P - 1, 0, st) == 0
    if T > 1:
        pad4 = torch.full((T * B, P + 4), 7.0, device="cuda")
        assert lib.torchain_to2d(x.data_ptr(), B, P, T, pad4.data_ptr(), P + 4, 0, st) == 1, _lib.last_error()
        assert torch.equal(pad4[:, :P], want) and bool((pad4[:, P:] == 7.0).all())
        back4 = torch.empty_like(x)
        assert lib.torchain_from2d(pad4.data_ptr(), P + 4, B, P, T, back4.data_ptr(), 0.5, 0, st) == 1, _lib.last_error()
        assert torch.equal(back4, 0.5 * x)


def test_gpu_supervision_merge_matches_appended_fst():
    """A minibatch assembled with Supervision.merge from separate examples gives the numbers of the FST
    Kaldi's AppendSupervision (Concat + RmEpsilon, synth.append_supervisions) would hand to the loss."""
    den = synth.make_den_fst(5, 120, 30, mean_deg=4.0)
    T = 7
    parts = [synth.make_supervision(40 + i, den, b, T, num_paths=3, weight=0.5) for i, b in enumerate((1, 1, 3, 2, 1))]
    B = sum(p.num_sequences for p in parts)
    appended = synth.append_supervisions(parts)
    assert appended.num_sequences == B
    y = synth.make_nnet_output(9, B, T, den.num_pdfs, scale=1.5)
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(appended), y, l2=0.01, leaky=0.1, mode="fp64", want_xent=True)
    g, s_ref = handles(den, appended)
    hs = [handles(den, p)[1] for p in parts]
    merged = io.Supervision.merge(hs)
    assert merged.shape == (B, T, den.num_pdfs) and merged.num_chains == B
    for s in (merged, s_ref):
        res, d, x = run_abi(g, s, y, 0.01, 0.1, xent=0.1, want_xent=True)
        assert rel(res[0], r["objf"]) < OBJF_RTOL
        assert rel(res[1], r["l2_term"]) < OBJF_RTOL
        assert res[2] == pytest.approx(r["weight"])
        assert deriv_err(d, r["deriv"]) < DERIV_RTOL
        assert deriv_err(x, r["xent"]) < DERIV_RTOL
    # merging is associative: ((0,1),(2,3,4)) == (0,1,2,3,4)
    nested = io.Supervision.merge([io.Supervision.merge(hs[:2]), io.Supervision.merge(hs[2:])])
    res2, d2, _ = run_abi(g, nested, y, 0.01, 0.1)
    assert rel(res2[0], r["objf"]) < OBJF_RTOL and deriv_err(d2, r["deriv"]) < DERIV_RTOL


def test_gpu_async_calls_calibrate_and_stay_correct():
    """Calls that never synchronise (results = NULL) still feed the partition autotuner: each leaves
    an event behind and a later call harvests the per-warp counters and re-deals the states.  Every
    call, before and after each re-partition and with back-to-back launches in flight, must match."""
    import torch
    lib = _lib.load()
    B, T, N, P = 128, 30, 3000, 300
    den = synth.make_den_fst(21, N, P, mean_deg=20.0)
    sup = synth.make_supervision(22, den, B, T, num_paths=2)
    y_np = synth.make_nnet_output(23, B, T, P, scale=1.5)
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y_np, l2=5e-5, leaky=0.1, mode="fp64", threads=8)
    g, s = handles(den, sup)
    y = torch.from_numpy(y_np).cuda()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = []
    for i in range(8):
        d = torch.full_like(y, float("nan"))
        res_dev = torch.zeros(3, device="cuda")
        ok = lib.torchain_ComputeChainObjfAndDeriv_ex(
            g.ptr, s.ptr, y.data_ptr(), y.shape[0], P, P, None, res_dev.data_ptr(), d.data_ptr(), P, None, 0,
            5e-5, 0.1, 0.0, 1.0, 0, stream)
        assert ok == 1, _lib.last_error()
        outs.append((res_dev, d))
        if i % 2 == 1:   # every other call is enqueued while the previous one is still running
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    for res_dev, d in outs:
        rr = res_dev.cpu().numpy()
        assert rel(rr[0], r["objf"]) < OBJF_RTOL and rel(rr[1], r["l2_term"]) < OBJF_RTOL
        assert deriv_err(d.cpu().numpy(), r["deriv"]) < DERIV_RTOL
    res, d, _ = run_abi(g, s, y_np, 5e-5, 0.1)   # and a synchronising call afterwards
    assert rel(res[0], r["objf"]) < OBJF_RTOL and deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_host_buffer_entry_padded_rows():
    """Host buffers whose row stride exceeds the row length (pitched 2-D copies; densely packed rows go as
    one linear copy): same results, and the padding of the derivative buffer is left alone."""
    lib = _lib.load()
    c = load_golden(GOLDEN[2])
    g, s = handles(c["den"], c["sup"])
    y = np.ascontiguousarray(c["y"], dtype=np.float32)
    rows, cols = y.shape
    yp = np.zeros((rows, cols + 5), dtype=np.float32)
    yp[:, :cols] = y
    dp = np.full((rows, cols + 3), np.nan, dtype=np.float32)
    res = np.zeros(3, dtype=np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, p(yp), rows, cols, cols + 5, p(res),
                                                    p(dp), cols + 3, None, 0, c["l2"], c["leaky"], 0.0, 0)
    assert ok == 1, _lib.last_error()
    assert rel(res[0], c["objf"]) < OBJF_RTOL
    assert deriv_err(dp[:, :cols], c["deriv"]) < DERIV_RTOL
    assert np.isnan(dp[:, cols:]).all()


def test_reference_training_loop():
    """The reference's integration test (test/test.py:19-78: Conv1d model with LF-MMI and xent heads, 11 SGD
    steps, l2 0.01, xent 0.01, kaldi_way=True, (B, C, T) model outputs handed to chain_loss as they are) on
    synthetic data, with the assertions tests/training_smoke.py adds; the CPU suite pins the same
    expectations with an oracle-backed loss (tests/test_oracle.py)."""
    import training_smoke as ts
    from torchain_b200.functions import chain_loss
    den, sup, x = ts.make_task()
    g, s = handles(den, sup)
    hist = ts.run_sgd(chain_loss, g, s, x, den.num_pdfs, "cuda")
    ts.check_history(hist, sup.num_sequences, sup.frames_per_sequence)


def test_host_buffer_entry_odd_row_bytes():
    """B*P not a multiple of 8 floats (B=3, P=61: 732 bytes per frame): 32-byte sectors straddle frame -- and chunk --
    boundaries, so while the forward launch reads frame t the copy engine may still be writing the other half of the
    same sector.  nnet_output is therefore read with ld.global.cg, never through the non-coherent path."""
    den = synth.make_den_fst(31, 300, 61, mean_deg=5.0)
    B, T = 3, 40
    sup = synth.make_supervision(32, den, B, T, num_paths=3, weight=1.0)
    y = synth.make_nnet_output(33, B, T, den.num_pdfs, scale=1.5)
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=5e-5, leaky=0.1, mode="fp64")
    g, s = handles(den, sup)
    for _ in range(5):
        res, d = _host_entry(g, s, y, 5e-5, 0.1)
        assert rel(res[0], r["objf"]) < OBJF_RTOL
        assert deriv_err(d, r["deriv"]) < DERIV_RTOL


def test_gpu_two_threads_one_stream():
    """Two host threads calling on the SAME (device, stream) -- the legacy default stream -- share one workspace
    (alpha store, zeroed scalars, barrier counters, pinned result words); the calls must serialise, not interleave."""
    import threading
    import torch
    lib = _lib.load()
    cases = []
    for k in range(2):
        den = synth.make_den_fst(40 + k, 400, 50, mean_deg=5.0) if k == 0 else cases[0][0]
        sup = synth.make_supervision(50 + k, den, 20 + 7 * k, 12, num_paths=3, weight=1.0)
        y = synth.make_nnet_output(60 + k, sup.num_sequences, 12, den.num_pdfs, scale=1.5)
        r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=5e-5, leaky=0.1, mode="fp64")
        cases.append((den, sup, y, r))
    den = cases[0][0]
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    sups = [io.Supervision.from_arrays(c[1].weight, c[1].num_sequences, 12, c[1].label_dim, c[1].arc_begin,
                                       c[1].arc_dst, c[1].arc_label, c[1].arc_cost, c[1].final_cost) for c in cases]
    ys = [torch.from_numpy(c[2]).cuda() for c in cases]
    ds = [torch.empty_like(y) for y in ys]
    torch.cuda.synchronize()
    errors = []

    def work(k):
        try:
            res = np.zeros(3, dtype=np.float32)
            for _ in range(30):
                ds[k].fill_(float("nan"))
                ok = lib.my_lib_ComputeChainObjfAndDeriv(
                    g.ptr, sups[k].ptr, ys[k].data_ptr(), ys[k].shape[0], ys[k].shape[1], ys[k].stride(0),
                    res.ctypes.data_as(ctypes.c_void_p), ds[k].data_ptr(), ds[k].stride(0), None, 0,
                    5e-5, 0.1, 0.0, 0, None)     # stream 0: the legacy default stream, both threads
                assert ok == 1, _lib.last_error()
                assert rel(res[0], cases[k][3]["objf"]) < OBJF_RTOL, (k, res)
                torch.cuda.synchronize()
                assert deriv_err(ds[k].cpu().numpy(), cases[k][3]["deriv"]) < DERIV_RTOL
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_gpu_fresh_supervision_per_step_prefetched_from_a_worker_thread():
    """A training loop builds a new Supervision for every minibatch
    (/root/reference/example/chime5/train_faster.py:116-141).  Here a worker thread builds and prefetches
    (asynchronous upload on its own stream) the next handle while the current step runs; handles die right after
    their step, their pooled device blocks are reused by later minibatches while kernels of earlier ones may still
    be running.  Every step must match the oracle."""
    import threading
    import torch
    from torchain_b200.functions import chain_loss
    den = synth.make_den_fst(71, 600, 80, mean_deg=6.0)
    B, T, n_steps = 24, 16, 12
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    oden = ol.OracleDen(den)
    sups = [synth.make_supervision(100 + k, den, B, T, num_paths=3, weight=1.0) for k in range(n_steps)]
    y = synth.make_nnet_output(72, B, T, den.num_pdfs, scale=1.5)
    refs = [ol.chain(oden, ol.OracleSup(sp), y, l2=5e-5, leaky=0.1, mode="fp64") for sp in sups]
    yt = torch.from_numpy(y).cuda()
    side = torch.cuda.Stream()
    box = {}

    def prepare(k):
        sp = sups[k]
        h = io.Supervision.from_arrays(sp.weight, B, T, sp.label_dim, sp.arc_begin, sp.arc_dst, sp.arc_label,
                                       sp.arc_cost, sp.final_cost)
        box[k] = h.prefetch(0, side)

    prepare(0)
    for k in range(n_steps):
        worker = threading.Thread(target=prepare, args=(k + 1,)) if k + 1 < n_steps else None
        if worker:
            worker.start()
        inp = yt.clone().requires_grad_(True)
        loss, results = chain_loss(inp, g, box.pop(k), l2_regularize=5e-5, leaky_hmm_coefficient=0.1)
        loss.backward()                      # the handle of step k is garbage from here on
        assert rel(float(results.data[0]), refs[k]["objf"]) < OBJF_RTOL
        assert deriv_err(inp.grad.cpu().numpy(), -refs[k]["deriv"]) < DERIV_RTOL
        if worker:
            worker.join()
    _lib.load().torchain_supervision_pool_trim()


def test_to2d_backward_accepts_any_gradient_layout():
    """a row-broadcast gradient (strides (0, 1)) or a non-float one reaches _To2d.backward."""
    import torch
    from torchain_b200.functions import to2d
    x = torch.randn(4, 33, 50, device="cuda", requires_grad=True)
    to2d(x).sum(0).sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    x.grad = None
    w = torch.randn(50 * 4, 33, device="cuda", dtype=torch.float64)
    (to2d(x).double() * w).sum().backward()
    ref = w.float().view(50, 4, 33).permute(1, 2, 0)
    assert torch.allclose(x.grad, ref)
