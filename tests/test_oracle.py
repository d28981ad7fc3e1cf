"""CPU suite, part 1: the oracle (oracle/chain_oracle.cpp) against the golden
fixtures, against the independent dense float64 formulation, and against the
property assertions of the reference's own vendored Kaldi test
(/root/reference/src/chain-supervision-test.hpp)."""
import numpy as np
import pytest

import dense_ref
import oracle_lib as ol
from common import DERIV_RTOL, GOLDEN, OBJF_RTOL, deriv_err, load_golden, rel
from torchain_b200 import synth


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p[-10:-4])
def test_oracle_fp64_matches_golden(path):
    g = load_golden(path)
    oden, osup = ol.OracleDen(g["den"]), ol.OracleSup(g["sup"])
    r = ol.chain(oden, osup, g["y"], l2=g["l2"], leaky=g["leaky"], mode="fp64")
    assert r["ok"]
    assert rel(r["objf"], g["objf"]) < 1e-7
    assert abs(r["l2_term"] - g["l2_term"]) <= 1e-9 * max(1.0, abs(g["l2_term"]))
    assert r["weight"] == g["weight"]
    assert deriv_err(r["deriv"], g["deriv"]) < 1e-7


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p[-10:-4])
def test_oracle_kaldi_fp32_within_stated_tolerance(path):
    g = load_golden(path)
    oden, osup = ol.OracleDen(g["den"]), ol.OracleSup(g["sup"])
    r = ol.chain(oden, osup, g["y"], l2=g["l2"], leaky=g["leaky"], mode="kaldi_fp32")
    assert rel(r["objf"], g["objf"]) < OBJF_RTOL
    assert abs(r["l2_term"] - g["l2_term"]) <= OBJF_RTOL * max(1e-30, abs(g["l2_term"]))
    assert deriv_err(r["deriv"], g["deriv"]) < DERIV_RTOL


def test_initial_probs_match_vectorised_restatement():
    for seed in range(5):
        den = synth.make_den_fst(seed, 50 + 37 * seed, 11, mean_deg=3.0)
        pi = ol.OracleDen(den).initial_probs()
        ref = synth.initial_probs(den)
        assert np.abs(pi - ref).max() < 1e-7
        assert abs(pi.sum() - 1.0) < 1e-5


@pytest.mark.parametrize("seed", range(12))
def test_oracle_vs_dense_autograd(seed):
    c = synth.small_case(seed)
    if c["T"] > 25:  # the dense reference is a python loop; keep the CPU suite fast
        c = dict(c)
        c["T"] = T = 7
        c["sup"] = synth.make_supervision(seed + 1, c["den"], c["B"], T, num_paths=3, weight=c["sup"].weight)
        c["y"] = synth.make_nnet_output(seed + 2, c["B"], T, c["P"])
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    r = ol.chain(oden, osup, c["y"], l2=c["l2"], leaky=c["leaky"], mode="fp64")
    objf, l2_term, w, g = dense_ref.chain_objf(c["den"], oden.initial_probs(), c["sup"], c["y"], c["leaky"], c["l2"])
    assert rel(r["objf"], objf) < 1e-7
    assert abs(r["l2_term"] - l2_term) <= 1e-9 * max(1.0, abs(l2_term))
    assert r["weight"] == pytest.approx(w)
    assert deriv_err(r["deriv"], g) < 1e-7


def test_batched_autograd_equals_per_sequence_autograd():
    c = synth.small_case(3)
    pi = ol.OracleDen(c["den"]).initial_probs()
    c = dict(c)
    c["T"] = T = 6
    c["sup"] = synth.make_supervision(4, c["den"], c["B"], T, num_paths=3, weight=c["sup"].weight)
    c["y"] = synth.make_nnet_output(5, c["B"], T, c["P"])
    a = dense_ref.chain_objf(c["den"], pi, c["sup"], c["y"], c["leaky"], c["l2"])
    b = dense_ref.chain_objf_batched(c["den"], pi, c["sup"], c["y"], c["leaky"], c["l2"])
    assert rel(a[0], b[0]) < 1e-12 and rel(a[1], b[1]) < 1e-12 and a[2] == b[2]
    assert deriv_err(a[3], b[3]) < 1e-12


@pytest.mark.parametrize("name,B,T", [("chime5", 4, 50), ("large_tree", 2, 20)])
def test_oracle_vs_independent_autograd_at_full_graph_size(name, B, T):
    """The oracle against the second formulation (dense_ref.den_logprob_batched: index_add recursion +
    autograd) on the den graphs of BASELINE.json cfg 2 (N=10 000, A~300k, all T=50 frames) and cfg 4:
    the sizes the GPU parity tests and bench.py's parity block trust the oracle at."""
    c = synth.CONFIGS[name]
    den = synth.make_den_fst(1234, c["N"], c["P"], mean_deg=c["mean_deg"])
    oden = ol.OracleDen(den)
    sup = synth.make_supervision(77, den, B, T, num_paths=4, weight=0.75)
    y = synth.make_nnet_output(78, B, T, den.num_pdfs, scale=2.0)
    r = ol.chain(oden, ol.OracleSup(sup), y, l2=c["l2"], leaky=c["leaky"], mode="fp64", threads=4)
    objf, l2_term, w, g = dense_ref.chain_objf_batched(den, oden.initial_probs(), sup, y, c["leaky"], c["l2"])
    assert r["ok"]
    assert rel(r["objf"], objf) < 1e-8
    assert rel(r["l2_term"], l2_term) < 1e-10
    assert r["weight"] == pytest.approx(w)
    assert deriv_err(r["deriv"], g) < 1e-7
    rk = ol.chain(oden, ol.OracleSup(sup), y, l2=c["l2"], leaky=c["leaky"], mode="kaldi_fp32", threads=4)
    assert rel(rk["objf"], objf) < OBJF_RTOL and deriv_err(rk["deriv"], g) < DERIV_RTOL


@pytest.mark.parametrize("seed", range(50))
def test_chain_training_properties(seed):
    """ChainTrainingTest assertions (chain-supervision-test.hpp:267-285)."""
    c = synth.small_case(seed)
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    r = ol.chain(oden, osup, c["y"], l2=0.0, leaky=c["leaky"], mode="kaldi_fp32")
    assert r["ok"]
    d = r["deriv"].astype(np.float64)
    assert np.linalg.norm(d.sum(axis=1)) < 0.1      # :267-273 shift invariance
    assert d.sum() < 0.2                            # :278-283
    assert r["objf"] <= 0.0                         # :285
    assert r["weight"] == pytest.approx(c["sup"].weight * c["B"] * c["T"])


@pytest.mark.parametrize("seed", range(8))
def test_chain_finite_difference(seed):
    """chain-supervision-test.hpp:287-340: predicted vs observed objf change within 25 %."""
    c = synth.small_case(seed)
    if c["T"] <= 2:
        pytest.skip("reference skips the comparison for very short sequences (:330)")
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    r = ol.chain(oden, osup, c["y"], l2=0.0, leaky=c["leaky"], mode="fp64")
    rng = np.random.default_rng(seed)
    pred, obs = [], []
    for _ in range(5):
        delta = rng.standard_normal(c["y"].shape) * 1e-4
        pred.append(float((r["deriv"] * delta).sum()))
        r2 = ol.chain(oden, osup, c["y"].astype(np.float64) + delta, l2=0.0, leaky=c["leaky"], mode="fp64",
                      want_deriv=False)
        obs.append(r2["objf"] - r["objf"])
    pred, obs = np.array(pred), np.array(obs)
    if np.linalg.norm(pred) > 0.1 * 1e-4:
        assert np.linalg.norm(pred - obs) <= 0.25 * np.linalg.norm(pred)


@pytest.mark.parametrize("seed", range(10))
def test_denominator_properties(seed):
    """ChainDenominatorTest (chain-supervision-test.hpp:388-463): occupancies sum to B*T,
    finite differences of the forward log-prob agree with the derivative."""
    c = synth.small_case(seed)
    oden = ol.OracleDen(c["den"])
    B, T = c["B"], c["T"]
    lp, d, ok = ol.den_only(oden, B, T, c["y"], leaky=1e-5)
    assert ok
    assert abs(d.sum() - B * T) < 1e-6 * B * T
    assert np.abs(d.reshape(T, B, -1).sum(axis=2) - 1.0).max() < 1e-9   # one unit of occupancy per (t, s)
    rng = np.random.default_rng(seed)
    for _ in range(3):
        delta = rng.standard_normal(c["y"].shape) * 1e-4
        lp2, _, _ = ol.den_only(oden, B, T, c["y"].astype(np.float64) + delta, leaky=1e-5, want_deriv=False)
        pred = float((d * delta).sum())
        if abs(pred) > 1e-7:
            assert abs((lp2 - lp) - pred) < 0.05 * abs(pred) + 1e-9


def test_numerator_shift_property():
    """TestSupervisionNumerator (:139-149): adding r_row to every element of a row changes
    the numerator log-prob by the sum of the r_row."""
    c = synth.small_case(3)
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    y = c["y"].astype(np.float64)
    r0 = ol.chain(oden, osup, y, leaky=0.1, mode="fp64", want_deriv=False)
    shift = np.random.default_rng(0).standard_normal(y.shape[0])
    r1 = ol.chain(oden, osup, y + shift[:, None], leaky=0.1, mode="fp64", want_deriv=False)
    assert abs((r1["num"] - r0["num"]) - c["sup"].weight * shift.sum()) < 1e-8 * max(1.0, abs(shift.sum()))


def test_guard_on_nan_input():
    """chain-training.cc guard: non-finite objf => objf = -10*weight, zero derivative (+ l2 part)."""
    c = synth.small_case(1)
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    y = c["y"].copy()
    y[0, 0] = np.nan
    r = ol.chain(oden, osup, y, l2=0.0, leaky=0.1, mode="kaldi_fp32")
    assert r["objf"] == pytest.approx(-10.0 * r["weight"])
    assert np.all(r["deriv"] == 0.0)


def test_threaded_oracle_is_bit_identical():
    c = synth.small_case(5)
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    a = ol.chain(oden, osup, c["y"], l2=0.01, leaky=0.1, mode="kaldi_fp32", threads=1)
    b = ol.chain(oden, osup, c["y"], l2=0.01, leaky=0.1, mode="kaldi_fp32", threads=4)
    assert a["objf"] == b["objf"] and np.array_equal(a["deriv"], b["deriv"])


def test_oracle_rejects_bad_inputs():
    c = synth.small_case(2)
    oden, osup = ol.OracleDen(c["den"]), ol.OracleSup(c["sup"])
    with pytest.raises(ValueError):
        ol.chain(oden, osup, c["y"][:-1], leaky=0.1)          # wrong number of rows
    with pytest.raises(ValueError):
        ol.chain(oden, osup, c["y"], leaky=0.0)               # Kaldi asserts 0 < leaky < 1


@pytest.mark.parametrize("kind", ["uniform", "tied", "mostly_tied"])
def test_state_splitting_by_incoming_pdf_preserves_the_loss(kind):
    """Groundwork for the tied-pdf recurrence: giving every state one copy per distinct
    incoming pdf (copies repeat the out-arcs) leaves objf and the derivative unchanged -- copies have identical
    futures and their forward mass adds up, Kaldi's initial-probability iteration included."""
    P = 30
    if kind == "uniform":
        den = synth.make_den_fst(11, 150, P, mean_deg=5.0)
    else:
        den = synth.make_den_fst_tied(11, 150, P, mean_deg=5.0, untied_frac=0.0 if kind == "tied" else 0.15)
    split, copy_of = synth.split_states_by_in_pdf(den)
    runs, with_in = synth.in_pdf_runs(den)
    assert split.num_states == runs + (den.num_states - with_in)
    if kind == "tied":
        assert runs <= 2 * den.num_states          # forward pdf + self-loop pdf
    # strictly tied now: one incoming pdf per state
    assert synth.in_pdf_runs(split)[0] == np.unique(split.dst).shape[0]
    assert np.array_equal(np.sort(np.unique(copy_of)), np.arange(den.num_states))
    B, T = 4, 7
    sup = synth.make_supervision(12, den, B, T, num_paths=3)
    y = synth.make_nnet_output(13, B, T, P, scale=1.5)
    a = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=0.0, leaky=0.1, mode="fp64")
    b = ol.chain(ol.OracleDen(split), ol.OracleSup(sup), y, l2=0.0, leaky=0.1, mode="fp64")
    assert abs(a["objf"] - b["objf"]) < 1e-8 * abs(a["objf"])      # initial probs are rounded to fp32 (Kaldi)
    assert np.abs(a["deriv"] - b["deriv"]).max() < 1e-7


def test_run_based_recurrence_matches_oracle():
    """tools/tied_pdf_study.py: the run-based ((state, incoming pdf)) form of the denominator recurrence --
    exp(y) once per run, one 4-byte gather per arc backward, one occupancy update per run -- is algebraically
    the per-arc recurrence: same log-probability, same occupancies, on tied and on uniform graphs."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "tied_pdf_study", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                                       "tied_pdf_study.py"))
    study = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(study)
    P, B, T = 25, 3, 6
    for den in (synth.make_den_fst(31, 120, P, mean_deg=5.0), synth.make_den_fst_tied(31, 120, P, mean_deg=5.0)):
        sup = synth.make_supervision(32, den, B, T, num_paths=2, weight=1.0)
        y = synth.make_nnet_output(33, B, T, P, scale=1.5)
        r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=0.0, leaky=0.1, mode="fp64", want_xent=True)
        logp, gamma, runs = study.tied_gamma(den, y, B, T, 0.1)
        gamma_den = (r["xent"] - r["deriv"]) / sup.weight      # deriv = w (gamma_num - gamma_den), xent = w gamma_num
        # both sides round the initial probabilities to fp32 (Kaldi) from independently computed doubles: 1 ulp
        assert np.abs(gamma - gamma_den).max() < 2e-7
        assert np.abs(gamma.sum(axis=1) - 1.0).max() < 1e-9


def test_reference_training_loop_with_oracle_backed_loss():
    """tests/training_smoke.py (the reference's test/test.py loop) driven by an oracle-backed autograd
    function on the CPU: pins the expectations the GPU suite then holds the CUDA path to
    (tests/test_gpu_parity.py::test_reference_training_loop)."""
    import torch
    from torch.autograd import Function
    import training_smoke as ts
    from torchain_b200.functions import ChainResults

    den, sup, x = ts.make_task()
    oden, osup = ol.OracleDen(den), ol.OracleSup(sup)
    B, T, P = sup.num_sequences, sup.frames_per_sequence, den.num_pdfs

    class OracleLoss(Function):
        @staticmethod
        def forward(ctx, inp, xent_inp, results, l2, leaky, xent_reg):
            y = inp.detach().permute(2, 0, 1).reshape(T * B, P).double().numpy()
            r = ol.chain(oden, osup, y, l2=l2, leaky=leaky, mode="fp64", want_xent=True)
            results.data[0], results.data[1], results.data[2] = r["objf"], r["l2_term"], r["weight"]
            back = lambda d: torch.from_numpy(d).view(T, B, P).permute(1, 2, 0).to(inp.dtype)
            ctx.g, ctx.gx = -back(r["deriv"]), -xent_reg * back(r["xent"])
            return inp.new_tensor([float(results.loss)])

        @staticmethod
        def backward(ctx, grad_output):
            return ctx.g, ctx.gx, None, None, None, None

    def loss_fn(pred, den_graph, supervision, l2_regularize=0.0, leaky_hmm_coefficient=1e-5,
                xent_regularize=0.0, xent_input=None, kaldi_way=False):
        assert kaldi_way and xent_input is not None
        results = ChainResults()
        return OracleLoss.apply(pred, xent_input, results, l2_regularize, leaky_hmm_coefficient,
                                xent_regularize), results

    hist = ts.run_sgd(loss_fn, None, None, x, P, "cpu")
    ts.check_history(hist, B, T)
