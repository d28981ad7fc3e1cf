"""What does reading the network output in place buy?  

Times, with CUDA events on config 2 (or --config / --batch / --frames):
  a. the C-ABI call on the reference's (T*B, P) matrix                     (what bench.py times)
  b. the C-ABI strided call on a (B, P, T) tensor, gradient written in (B, P, T)
  c. chain_loss(to2d(x)) + backward(): the reference's route, permute+contiguous copy in,
     permuted gradient copy out (torchain/functions.py:118-138)
  d. chain_loss(x) + backward() on the (B, P, T) tensor: the library's tiled relayout kernels
  e. chain_loss on a (B, T, P) tensor viewed as (B, P, T): read and written in place
  f. the relayout kernels alone against torch's permute().contiguous()
and the DRAM-side consistency of b. against a. (same results, same gradient).
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from torchain_b200 import _lib, io  # noqa: E402
from torchain_b200.functions import chain_loss, to2d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="chime5")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--frames", type=int, default=0)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()

lib = _lib.load()
c = bench.make_workload(args.config, args.batch, args.frames, 0)
den, sup = c["den"], c["sup"]
B, T, P = c["B"], c["T"], c["P"]
g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start, den.final_cost,
                                    den.num_pdfs)
s = io.Supervision.from_arrays(sup.weight, B, T, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                               sup.arc_cost, sup.final_cost)
gen = torch.Generator(device="cuda").manual_seed(1234)
y2 = torch.randn(T * B, P, device="cuda", generator=gen) * 2.0
y3 = y2.view(T, B, P).permute(1, 2, 0).contiguous()  # (B, P, T)
d2, d3 = torch.empty_like(y2), torch.empty_like(y3)
res = np.zeros(3, dtype=np.float32)
rp = res.ctypes.data_as(ctypes.c_void_p)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
st2 = (ctypes.c_int64 * 3)(B * P, P, 1)
st3 = (ctypes.c_int64 * 3)(1, P * T, T)


def abi2d():
    assert lib.torchain_ComputeChainObjfAndDeriv_ex(g.ptr, s.ptr, y2.data_ptr(), T * B, P, P, rp, None, d2.data_ptr(), P,
                                                    None, 0, c["l2"], c["leaky"], 0.0, -1.0, 0, stream) == 1


def abi3d():
    assert lib.torchain_ComputeChainObjfAndDeriv_strided(g.ptr, s.ptr, y3.data_ptr(), T, B, P, st3, rp, None,
                                                         d3.data_ptr(), st3, None, None, c["l2"], c["leaky"], 0.0, -1.0,
                                                         0, stream) == 1, _lib.last_error()


def torch_to2d(x):  # the reference's to2d, verbatim semantics (functions.py:118-125)
    return x.permute(2, 0, 1).contiguous().view(-1, x.shape[1])


def op(route):
    x = (y3bt.permute(0, 2, 1) if route == "btp_view" else y3).detach().requires_grad_(True)
    loss, _ = chain_loss(torch_to2d(x) if route == "to2d" else x, g, s, l2_regularize=c["l2"],
                         leaky_hmm_coefficient=c["leaky"])
    loss.backward()
    return x.grad


y3bt = y3.permute(0, 2, 1).contiguous()  # (B, T, P): pdf innermost


def timeit(fn, n):
    for _ in range(4):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(4):  # partition autotuner rounds
    abi2d()
out = {}
abi2d()
r2 = res.copy()
abi3d()
r3 = res.copy()
torch.cuda.synchronize()
d3_as2 = d3.permute(2, 0, 1).reshape(T * B, P)
err = float((d3_as2 - d2).abs().max() / d2.abs().max())
print("results 2-D %s  strided %s  max |dgrad| / max |grad| = %.2e" % (r2, r3, err))
assert np.allclose(r2, r3, rtol=1e-5) and err < 1e-5
out["abi_2d_ms"] = timeit(abi2d, args.iters)
out["abi_bpt_ms"] = timeit(abi3d, args.iters)
out["op_to2d_ms"] = timeit(lambda: op("to2d"), args.iters)
out["op_relayout_ms"] = timeit(lambda: op("relayout"), args.iters)
out["op_btp_view_ms"] = timeit(lambda: op("btp_view"), args.iters)
rows = torch.empty_like(y2)
out["to2d_kernel_ms"] = timeit(lambda: lib.torchain_to2d(y3.data_ptr(), B, P, T, rows.data_ptr(), P, 0, stream), args.iters)
out["from2d_kernel_ms"] = timeit(lambda: lib.torchain_from2d(rows.data_ptr(), P, B, P, T, d3.data_ptr(), 1.0, 0, stream),
                                 args.iters)
out["torch_permute_ms"] = timeit(lambda: torch_to2d(y3), args.iters)
ga, gb, gc = op("to2d"), op("relayout"), op("btp_view")
assert float((ga - gb).abs().max() / ga.abs().max()) < 1e-5
assert float((ga - gc).abs().max() / ga.abs().max()) < 1e-5
nbytes = 2 * y2.numel() * 4
print("to2d kernel %.1f us (%.2f TB/s), from2d kernel %.1f us (%.2f TB/s), torch permute+contiguous %.1f us (%.2f TB/s)" % (
    out["to2d_kernel_ms"] * 1e3, nbytes / out["to2d_kernel_ms"] / 1e9, out["from2d_kernel_ms"] * 1e3,
    nbytes / out["from2d_kernel_ms"] / 1e9, out["torch_permute_ms"] * 1e3, nbytes / out["torch_permute_ms"] / 1e9))
print("C ABI, (T*B, P) matrix                         : %.3f ms" % out["abi_2d_ms"])
print("C ABI strided, (B, P, T) in place               : %.3f ms" % out["abi_bpt_ms"])
print("chain_loss(to2d(x)).backward()  [reference route]: %.3f ms" % out["op_to2d_ms"])
print("chain_loss(x).backward()  [library relayout]     : %.3f ms" % out["op_relayout_ms"])
print("chain_loss((B,T,P) viewed (B,P,T)).backward() [in place]: %.3f ms" % out["op_btp_view_ms"])
import json  # noqa: E402
print(json.dumps({"config": args.config, "B": B, "T": T, "P": P, **out}))
