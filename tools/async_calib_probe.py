"""Does a caller that never synchronises (results = NULL) get the autotuned partition?
Times config 2 with asynchronous calls only, with and without the autotuner (TORCHAIN_NO_AUTOTUNE=1)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import bench
    from torchain_b200 import _lib, io
    lib = _lib.load()
    c = bench.make_workload("chime5", 0, 0, 0)
    den, sup = c["den"], c["sup"]
    B, T, P = c["B"], c["T"], c["P"]
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    s = io.Supervision.from_arrays(sup.weight, B, T, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                                   sup.arc_cost, sup.final_cost)
    y = torch.randn(T * B, P, device="cuda") * 2.0
    d = torch.empty_like(y)
    res_dev = torch.zeros(3, device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        assert lib.torchain_ComputeChainObjfAndDeriv_ex(g.ptr, s.ptr, y.data_ptr(), T * B, P, P, None, res_dev.data_ptr(),
                                                        d.data_ptr(), P, None, 0, c["l2"], c["leaky"], 0.0, -1.0, 0,
                                                        stream) == 1
    for i in range(12):   # a training loop reads its loss every few steps
        step()
        if i % 2 == 1:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("%.3f" % (e0.elapsed_time(e1) / 20))
else:
    for label, env in (("autotuner on ", {}), ("autotuner off", {"TORCHAIN_NO_AUTOTUNE": "1"})):
        out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
        print("asynchronous calls only, %s: %s ms per step" % (label, out.stdout.strip() or out.stderr[-300:]))
