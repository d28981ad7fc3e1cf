"""Shared helpers for the test-suite."""
import glob
import os

import numpy as np

from torchain_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "case_*.npz")))

# Stated tolerances (BASELINE.json north_star): objf and l2 within 1e-5 relative,
# derivatives within 1e-4 of the largest reference derivative magnitude.
OBJF_RTOL = 1e-5
DERIV_RTOL = 1e-4


def load_golden(path):
    z = np.load(path)
    den = synth.DenFst(int(z["num_states"]), int(z["num_pdfs"]), int(z["den_start"]), z["den_src"], z["den_dst"],
                       z["den_ilabel"], z["den_cost"], z["den_final"])
    sup = synth.SupFst(float(z["sup_weight"]), int(z["B"]), int(z["T"]), int(z["num_pdfs"]), z["sup_arc_begin"],
                       z["sup_arc_dst"], z["sup_arc_label"], z["sup_arc_cost"], z["sup_final"])
    return dict(den=den, sup=sup, y=z["y"], leaky=float(z["leaky"]), l2=float(z["l2"]), objf=float(z["objf"]),
                l2_term=float(z["l2_term"]), weight=float(z["weight"]), deriv=z["deriv"])


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def deriv_err(d, ref):
    return float(np.abs(np.asarray(d, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))
