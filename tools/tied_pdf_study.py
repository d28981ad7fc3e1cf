"""Groundwork for the tied-pdf ("run-based") denominator recurrence of DESIGN.md section 8, CPU only.

A *run* is a (state, incoming pdf) pair.  In a real den.fst the pdf of an arc belongs to the HMM state being
entered, so a state has ~2 runs (its forward pdf and its self-loop pdf) for ~30 in-arcs.  Then, per frame,
  forward :  a_{t+1}(run) = e_t(pdf_run) / S_t * sum_{arcs i->j in run} alpha'_t(i) p        (exp(y) once per run)
             alpha_{t+1}(j) = sum_{runs of j} a_{t+1}(run)
  backward:  W(run) = (beta'_{t+1}(j) + b_{t+1}) * e_t(pdf_run)                               (once per run)
             beta'_t(h) = sum_{arcs h->j} p * W(run of the arc) / S_t                          (ONE 4-byte gather per arc)
             gamma_t(pdf_run) += a_{t+1}(run) * (beta'_{t+1}(j) + b_{t+1})                     (one RED per run, not per arc)
`tied_gamma` below is that recurrence in numpy (float64, one sequence at a time); `python tools/tied_pdf_study.py`
checks it against the oracle (Kaldi's per-arc formulation) and prints the run statistics and the L2-traffic model
for the synthetic graphs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from torchain_b200 import synth  # noqa: E402


def tied_gamma(den, y, B, T, leaky):
    """Denominator log-probability and occupancies gamma[t*B+s, pdf] by the run-based recurrence."""
    N, P = den.num_states, den.num_pdfs
    src, dst = den.src.astype(np.int64), den.dst.astype(np.int64)
    pdf = den.ilabel.astype(np.int64) - 1
    p = np.exp(-den.cost.astype(np.float64)).astype(np.float32).astype(np.float64)   # Kaldi keeps fp32 probs
    pi = synth.initial_probs(den).astype(np.float32).astype(np.float64)             # and fp32 initial probs
    key = dst * P + pdf
    uniq, run = np.unique(key, return_inverse=True)
    run_state, run_pdf = uniq // P, uniq % P
    R = uniq.shape[0]
    gamma = np.zeros((T * B, P))
    logp = 0.0
    for s in range(B):
        e = np.exp(y[s::B].astype(np.float64))            # [T, P]
        alpha = np.zeros((T + 1, N))
        a_run = np.zeros((T + 1, R))
        S = np.zeros(T + 1)
        alpha[0] = pi
        for t in range(T):
            S[t] = alpha[t].sum()
            mixed = alpha[t] + leaky * pi * S[t]
            part = np.bincount(run, weights=mixed[src] * p, minlength=R)
            a_run[t + 1] = e[t, run_pdf] * part / S[t]
            alpha[t + 1] = np.bincount(run_state, weights=a_run[t + 1], minlength=N)
        S[T] = alpha[T].sum()
        tot = (alpha[T] + leaky * pi * S[T]).sum()
        logp += np.log(tot) + np.log(S[:T]).sum()
        beta = np.full(N, 1.0 / tot)
        for t in range(T - 1, -1, -1):
            b = leaky * (pi * beta).sum()
            full = beta[run_state] + b                       # per run: beta'_{t+1}(j) + b_{t+1}
            np.add.at(gamma[t * B + s], run_pdf, a_run[t + 1] * full)
            W = full * e[t, run_pdf]
            beta = np.bincount(src, weights=p * W[run], minlength=N) / S[t]
    return logp, gamma, R


def main():
    import oracle_lib as ol
    P, B, T = 40, 3, 8
    for name, den in (("uniform pdfs", synth.make_den_fst(5, 300, P, mean_deg=8.0)),
                      ("tied pdfs", synth.make_den_fst_tied(5, 300, P, mean_deg=8.0)),
                      ("tied, 10 % odd arcs", synth.make_den_fst_tied(5, 300, P, mean_deg=8.0, untied_frac=0.1))):
        sup = synth.make_supervision(6, den, B, T, num_paths=3)
        y = synth.make_nnet_output(7, B, T, P, scale=1.5)
        r = ol.chain(ol.OracleDen(den), ol.OracleSup(sup), y, l2=0.0, leaky=0.1, mode="fp64", want_xent=True)
        logp, gamma, R = tied_gamma(den, y, B, T, 0.1)
        gamma_den_oracle = (r["xent"] - r["deriv"]) / sup.weight
        err = np.abs(gamma - gamma_den_oracle).max()
        A, N = den.src.shape[0], den.num_states
        now, runs = 20 * A, 8 * A + 3 * 4 * R   # bytes per frame and sequence through L2, see DESIGN.md section 8
        print("%-20s N=%d A=%d runs=%d (%.2f per state)  max |gamma - oracle| = %.1e  row sums in [%.6f, %.6f]  "
              "L2 bytes/frame/seq: %d per-arc formulation, %d run-based (%.2fx)" % (
                  name, N, A, R, R / N, err, gamma.sum(1).min(), gamma.sum(1).max(), now, runs, now / runs))
        assert err < 1e-6   # the two sides round pi to fp32 independently: up to ~1e-8
    for name, den in (("config 2 graph (uniform)", synth.make_den_fst(1234, 10000, 2928, mean_deg=28.0)),
                      ("config 2 shape, tied", synth.make_den_fst_tied(1234, 10000, 2928, mean_deg=28.0))):
        runs, _ = synth.in_pdf_runs(den)
        A = den.src.shape[0]
        print("%-26s A=%d runs=%d  L2 bytes/frame/seq: %d -> %d (%.2fx)" % (name, A, runs, 20 * A, 8 * A + 12 * runs,
                                                                            20 * A / (8 * A + 12 * runs)))


if __name__ == "__main__":
    main()
