"""ctypes loader of tools/microbench/libtorchain_microbench.so (bench-only kernels: the measured
roofline denominators and feasibility probes).  Never imported by torchain_b200/."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtorchain_microbench.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    lib = ctypes.CDLL(LIB_PATH)
    i, i64, d, p = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    for name, args in {
        "tb_microbench_l2_gather": [i, i64, i, i],
        "tb_microbench_l2_red": [i, i64, i, i],
        "tb_microbench_grid_barrier": [i, i],
        "tb_microbench_pattern": [i, i, i, i],
        "tb_microbench_pattern_async": [i, i, i, i, i],
        "tb_microbench_dsmem_pattern": [i, i, i, i, i, i, p],
        "tb_microbench_sliced": [i, i, i, i, i],
    }.items():
        fn = getattr(lib, name)
        fn.restype = d
        fn.argtypes = args
    lib.tb_microbench_launch_count.restype = ctypes.c_longlong
    _lib = lib
    return lib
