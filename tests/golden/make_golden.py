"""Generates tests/golden/case_*.npz.

The reference cannot be imported or built in this image (its arithmetic is in
Kaldi, absent) and ships no golden vectors for the chain
loss, so these fixtures are produced by tests/dense_ref.py: a dense-matrix
float64 formulation whose derivatives come from torch autograd, independent of
both oracle/ and the CUDA kernels.  They pin the oracle (CPU suite) and the
kernels (GPU suite) to the same numbers.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import dense_ref  # noqa: E402
from torchain_b200 import synth  # noqa: E402

CASES = [
    # (name, seed, N, P, B, T, K, mean_deg, leaky, l2, weight, y_scale, boundary_costs)
    ("a", 101, 12, 7, 2, 5, 2, 2.0, 1e-5, 0.0, 1.0, 1.0, False),
    ("b", 102, 40, 19, 3, 9, 3, 3.0, 0.2, 0.01, 0.5, 1.0, True),
    ("c", 103, 97, 33, 5, 12, 4, 4.0, 0.1, 5e-5, 1.0, 2.0, True),
    ("d", 104, 8, 5, 1, 2, 1, 1.0, 0.1, 0.0, 1.0, 0.0, False),   # all-zero nnet output, T=2, B=1
    ("e", 105, 150, 60, 4, 20, 4, 5.0, 1e-5, 0.01, 0.5, 1.5, False),
]


def main():
    for (name, seed, N, P, B, T, K, deg, leaky, l2, weight, ys, bc) in CASES:
        den = synth.make_den_fst(seed, N, P, mean_deg=deg)
        sup = synth.make_supervision(seed + 1, den, B, T, num_paths=K, weight=weight, boundary_final_costs=bc)
        y = synth.make_nnet_output(seed + 2, B, T, P, scale=ys, zero=(ys == 0.0))
        pi = synth.initial_probs(den).astype(np.float32)  # Kaldi keeps initial probs as float
        objf, l2_term, w, deriv = dense_ref.chain_objf(den, pi, sup, y, leaky, l2)
        np.savez_compressed(
            os.path.join(HERE, "case_%s.npz" % name),
            den_src=den.src, den_dst=den.dst, den_ilabel=den.ilabel, den_cost=den.cost,
            den_final=den.final_cost, den_start=den.start, num_states=N, num_pdfs=P,
            sup_arc_begin=sup.arc_begin, sup_arc_dst=sup.arc_dst, sup_arc_label=sup.arc_label,
            sup_arc_cost=sup.arc_cost, sup_final=sup.final_cost, sup_weight=weight, B=B, T=T,
            y=y, leaky=leaky, l2=l2, objf=objf, l2_term=l2_term, weight=w, deriv=deriv)
        print(name, objf, l2_term, w, float(np.abs(deriv).max()))


if __name__ == "__main__":
    main()
