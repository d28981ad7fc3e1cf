"""ctypes loader for libtorchain_b200.so (the C ABI in include/torchain_b200.h).

Replaces the reference's cffi module ``torchain._ext.my_lib`` built by
/root/reference/build.py:19-31 from /root/reference/src/my_lib.h.  The library
is built in-tree by ``torchain_b200/csrc/Makefile`` (or ``__graft_entry__.build()``).
There is no Python/CPU fallback: if the shared object is missing, importing any
compute entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TORCHAIN_B200_LIB", os.path.join(_HERE, "libtorchain_b200.so"))

_c_f = ctypes.c_float
_c_i = ctypes.c_int
_c_i32 = ctypes.c_int32
_c_i64 = ctypes.c_int64
_c_p = ctypes.c_void_p
_c_d = ctypes.c_double

# name -> (restype, argtypes); every symbol declared in include/torchain_b200.h
SIGNATURES = {
    "torchain_last_error": (ctypes.c_char_p, []),
    "torchain_abi_version": (_c_i, []),
    "torchain_kernel_launch_count": (_c_i64, []),
    "my_lib_set_kaldi_device": (_c_i, [_c_i]),
    "my_lib_denominator_graph_new": (_c_p, [ctypes.c_char_p, _c_i]),
    "torchain_denominator_graph_from_arrays": (
        _c_p, [_c_i32, _c_i32, _c_p, _c_p, _c_p, _c_p, _c_i32, _c_p, _c_i32]),
    "my_lib_denominator_graph_free": (None, [_c_p]),
    "torchain_denominator_graph_num_states": (_c_i, [_c_p]),
    "torchain_denominator_graph_num_arcs": (_c_i, [_c_p]),
    "torchain_denominator_graph_num_pdfs": (_c_i, [_c_p]),
    "torchain_denominator_graph_num_runs": (_c_i, [_c_p]),
    "torchain_denominator_graph_is_tied": (_c_i, [_c_p]),
    "torchain_denominator_graph_initial_probs": (_c_i, [_c_p, _c_p]),
    "torchain_supervision_from_arrays": (
        _c_p, [_c_f, _c_i32, _c_i32, _c_i32, _c_i32, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "torchain_supervision_from_kaldi_bytes": (_c_p, [_c_p, _c_i64, _c_p]),
    "torchain_supervision_from_chain_egs": (_c_p, [_c_p, _c_i64, _c_p]),
    "torchain_supervision_merge": (_c_p, [_c_p, _c_i32]),
    "torchain_supervision_prefetch": (_c_i, [_c_p, _c_i, _c_p]),
    "torchain_supervision_pool_trim": (None, []),
    "my_lib_supervision_new": (_c_p, [_c_p]),
    "my_lib_supervision_rand_new": (_c_p, [_c_p]),
    "my_lib_supervision_free": (None, [_c_p]),
    "my_lib_supervision_num_pdf": (_c_i, [_c_p]),
    "my_lib_supervision_num_sequence": (_c_i, [_c_p]),
    "my_lib_supervision_num_frame": (_c_i, [_c_p]),
    "torchain_supervision_num_chains": (_c_i, [_c_p]),
    "my_lib_ComputeChainObjfAndDeriv": (
        _c_i, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_p, _c_p, _c_i64, _c_p, _c_i64,
               _c_f, _c_f, _c_f, _c_i, _c_p]),
    "torchain_ComputeChainObjfAndDeriv_ex": (
        _c_i, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_i64,
               _c_f, _c_f, _c_f, _c_f, _c_i, _c_p]),
    "torchain_ComputeChainObjfAndDeriv_strided": (
        _c_i, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p,
               _c_f, _c_f, _c_f, _c_f, _c_i, _c_p]),
    "torchain_to2d": (_c_i, [_c_p, _c_i64, _c_i64, _c_i64, _c_p, _c_i64, _c_i, _c_p]),
    "torchain_from2d": (_c_i, [_c_p, _c_i64, _c_i64, _c_i64, _c_i64, _c_p, _c_f, _c_i, _c_p]),
    "torchain_ComputeChainObjfAndDeriv_host": (
        _c_i, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_i64, _c_p, _c_p, _c_i64, _c_p, _c_i64,
               _c_f, _c_f, _c_f, _c_i]),
    "torchain_host_chunk_bounds": (_c_i, [_c_i, _c_p, _c_i]),
    "torchain_profile_enable": (None, [_c_i]),
    "torchain_profile_read": (_c_i, [_c_p, _c_p, _c_p, _c_i]),
    "torchain_profile_timeline": (_c_i, [_c_p, _c_p, _c_i]),
    "torchain_profile_cta_trace": (_c_i64, [_c_p, _c_p, _c_i64, _c_p, _c_p]),
}

_lib = None


def load():
    """Load the shared library (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libtorchain_b200.so is not built (%s); run `make -C torchain_b200/csrc` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    msg = load().torchain_last_error()
    return msg.decode() if msg else ""
