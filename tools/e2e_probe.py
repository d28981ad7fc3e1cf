"""Where does the host-buffer entry spend its time?  Compares: device-resident call, host call
without derivative output, host call with zero-copy (pinned) derivative, host call with pageable
derivative (staged copy)."""
import ctypes, os, sys, time
import numpy as np
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
y = torch.randn(T * B, P, device="cuda") * 2.0
y_pin = torch.empty(T * B, P, pin_memory=True); y_pin.copy_(y)
d_pin = torch.empty(T * B, P, pin_memory=True)
d_page = np.empty((T * B, P), dtype=np.float32)
res = np.zeros(3, dtype=np.float32)
rp = res.ctypes.data_as(ctypes.c_void_p)

def host(dptr):
    ok = lib.torchain_ComputeChainObjfAndDeriv_host(g.ptr, s.ptr, y_pin.data_ptr(), T * B, P, P, rp, dptr, P, None, 0,
                                                    c["l2"], c["leaky"], 0.0, 0)
    assert ok == 1, _lib.last_error()

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

print("host, no derivative output : %.3f ms" % timeit(lambda: host(None)))
print("host, pinned (zero-copy)    : %.3f ms" % timeit(lambda: host(d_pin.data_ptr())))
print("host, pageable (staged copy): %.3f ms" % timeit(lambda: host(d_page.ctypes.data_as(ctypes.c_void_p))))
# raw PCIe numbers for context
t = timeit(lambda: y.copy_(y_pin, non_blocking=True)); print("H2D 75 MB pinned: %.3f ms (%.1f GB/s)" % (t, 75e-3 / t * 1e3 / 1e3 * 1e3 / 1e3 if False else 74.9568 / t))
t = timeit(lambda: d_pin.copy_(y, non_blocking=True)); print("D2H 75 MB pinned: %.3f ms (%.1f GB/s)" % (t, 74.9568 / t))
