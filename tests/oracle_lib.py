"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module (see oracle/chain_oracle.cpp header).
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(ORACLE_DIR, "chain_oracle.cpp")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        build()
    lib = ctypes.CDLL(LIB)
    P, I, I64, F, D = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double
    lib.oracle_last_error.restype = ctypes.c_char_p
    lib.oracle_max_threads.restype = ctypes.c_int
    lib.oracle_den_graph_new.restype = P
    lib.oracle_den_graph_new.argtypes = [I, I, P, P, P, P, I, P, I]
    lib.oracle_den_graph_free.argtypes = [P]
    lib.oracle_den_graph_initial_probs.argtypes = [P, P]
    lib.oracle_supervision_new.restype = P
    lib.oracle_supervision_new.argtypes = [F, I, I, I, I, P, P, P, P, P]
    lib.oracle_supervision_free.argtypes = [P]
    for name in ("oracle_chain_f32", "oracle_chain_f64"):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [P, P, P, I64, I64, I64, D, D, ctypes.c_int, ctypes.c_int, P, P, I64, P, I64]
    lib.oracle_den_only_f64.restype = ctypes.c_int
    lib.oracle_den_only_f64.argtypes = [P, I, I, P, I64, D, P, P, I64]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleDen:
    def __init__(self, den):
        lib = load()
        self._keep = [np.ascontiguousarray(x) for x in (den.src, den.dst, den.ilabel, den.cost, den.final_cost)]
        s, d, il, c, f = self._keep
        self.N, self.P = den.num_states, den.num_pdfs
        self.ptr = lib.oracle_den_graph_new(den.num_states, den.num_arcs, _p(s), _p(d), _p(il), _p(c),
                                            den.start, _p(f), den.num_pdfs)
        if not self.ptr:
            raise ValueError(lib.oracle_last_error().decode())

    def initial_probs(self):
        out = np.empty(self.N, dtype=np.float32)
        load().oracle_den_graph_initial_probs(self.ptr, _p(out))
        return out

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr and _lib is not None:
            try:
                _lib.oracle_den_graph_free(ptr)
            except Exception:  # interpreter shutdown
                pass


class OracleSup:
    def __init__(self, sup):
        lib = load()
        self._keep = [np.ascontiguousarray(x) for x in
                      (sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)]
        ab, ad, al, ac, fc = self._keep
        self.B, self.T = sup.num_sequences, sup.frames_per_sequence
        self.ptr = lib.oracle_supervision_new(sup.weight, sup.num_sequences, sup.frames_per_sequence,
                                              sup.label_dim, sup.num_states, _p(ab), _p(ad), _p(al), _p(ac), _p(fc))
        if not self.ptr:
            raise ValueError(lib.oracle_last_error().decode())

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr and _lib is not None:
            try:
                _lib.oracle_supervision_free(ptr)
            except Exception:
                pass


def chain(oden, osup, y, l2=0.0, leaky=1e-5, mode="fp64", threads=1, want_deriv=True, want_xent=False,
          clamp_exp=False):
    """Returns dict(objf, l2_term, weight, num, den, ok, deriv, xent)."""
    lib = load()
    dt = np.float64 if mode == "fp64" else np.float32
    y = np.ascontiguousarray(y, dtype=dt)
    rows, cols = y.shape
    res = np.zeros(6, dtype=np.float64)
    deriv = np.empty_like(y) if want_deriv else None
    xent = np.empty_like(y) if want_xent else None
    fn = lib.oracle_chain_f64 if mode == "fp64" else lib.oracle_chain_f32
    ok = fn(oden.ptr, osup.ptr, _p(y), rows, cols, cols, float(l2), float(leaky), int(threads), int(clamp_exp),
            _p(res), _p(deriv), cols, _p(xent), cols)
    if ok != 1:
        raise ValueError(lib.oracle_last_error().decode())
    return dict(objf=res[0], l2_term=res[1], weight=res[2], num=res[3], den=res[4], ok=bool(res[5]),
                deriv=deriv, xent=xent)


def den_only(oden, B, T, y, leaky=1e-5, want_deriv=True):
    lib = load()
    y = np.ascontiguousarray(y, dtype=np.float64)
    lp = ctypes.c_double(0.0)
    deriv = np.empty_like(y) if want_deriv else None
    ok = lib.oracle_den_only_f64(oden.ptr, B, T, _p(y), y.shape[1], float(leaky), ctypes.byref(lp), _p(deriv),
                                 y.shape[1])
    return lp.value, deriv, bool(ok)
