"""Host-side timeline of the host-buffer entry (TORCHAIN_HOST_TRACE=1) on BASELINE.json config 2.  `end` of a trace
line is the duration of the call.  (Round 2 measured the experiment variants of the entry with this tool --
profiles/r02_host_trace_variants.txt -- kept two copy streams and deleted the rest.)"""
import ctypes, os, sys
import numpy as np
os.environ["TORCHAIN_HOST_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from torchain_b200 import _lib, io

lib = _lib.load()
c = bench.make_workload("chime5", 0, 0, 0)
den, sup = c["den"], c["sup"]
B, T, P = c["B"], c["T"], c["P"]
g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start, den.final_cost, den.num_pdfs)
s = io.Supervision.from_arrays(sup.weight, B, T, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)
y_pin = torch.empty(T * B, P, pin_memory=True)
y_pin.copy_(torch.randn(T * B, P) * 2.0)
d_pin = torch.empty(T * B, P, pin_memory=True)
res = np.zeros(3, dtype=np.float32)
rp = res.ctypes.data_as(ctypes.c_void_p)


def host():
    ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, y_pin.data_ptr(), T * B, P, P, rp, d_pin.data_ptr(), P,
                                                    None, 0, c["l2"], c["leaky"], 0.0, 0)
    assert ok == 1, _lib.last_error()


def variant(name, env, n=6):
    for k, v in env.items():
        os.environ[k] = v
    sys.stderr.write("== %s\n" % name)
    sys.stderr.flush()
    for _ in range(n):
        host()
    for k in env:
        del os.environ[k]


variant("warm-up (partition calibration)", {}, 5)
variant("default (chunks alternate between two copy streams, progress word released at system scope)", {}, 8)
# results must not depend on the variant
print("objf %.4f l2_term %.4f weight %.1f  deriv finite: %s" % (res[0], res[1], res[2], bool(torch.isfinite(d_pin).all())))
