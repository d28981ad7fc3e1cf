"""``torchain._ext.my_lib`` -- the object the reference obtained from its cffi build
(/root/reference/build.py:19-31, declarations /root/reference/src/my_lib.h), here a thin adapter
over the C ABI of libtorchain_b200.so (include/torchain_b200.h, bound in torchain_b200/_lib.py).

The cffi module converted torch tensors to TH*Tensor* on the way in; this adapter does the same
conversion to (device pointer, rows, cols, row stride, device index, current stream), so reference
code such as ::

    from torchain._ext import my_lib
    ffi = my_lib._ffi
    my_lib.my_lib_ComputeChainObjfAndDeriv(den_graph, supervision, input, results.data,
                                           mmi_grad, ffi.NULL, l2, leaky, xent)

(/root/reference/torchain/functions.py:75-86) runs unmodified.  Only the chain-loss path is
bound; the Kaldi egs readers (my_lib_example_*) are outside the hot path and raise.
"""
import ctypes

import torch

from torchain_b200 import _lib


class _Null:
    """ffi.NULL"""

    def __bool__(self):
        return False

    def __eq__(self, other):
        return other is None or isinstance(other, _Null) or other == 0

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return 0

    def __repr__(self):
        return "<cdata 'void *' NULL>"


class _FFI:
    """The two things the reference uses of cffi's ffi object: ``NULL`` and ``new("char[]", bytes)``."""
    NULL = _Null()

    @staticmethod
    def new(ctype, init=None):
        if ctype != "char[]":
            raise TypeError("only ffi.new('char[]', bytes) is supported")
        return ctypes.create_string_buffer(init if init is not None else b"")


def _is_null(x):
    return x is None or isinstance(x, _Null)


def _ptr(handle):
    return None if _is_null(handle) else handle


def _matrix(t, what):
    if t.dim() != 2:
        raise RuntimeError("dim != 2")            # /root/reference/src/common.hpp:110
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError(what + " must be a CUDA float tensor")
    if t.stride(1) != 1:
        raise RuntimeError(what + " rows must be dense")
    return t.data_ptr(), t.shape[0], t.shape[1], t.stride(0)


class _MyLib:
    _ffi = _FFI()

    def __init__(self):
        self._my_lib = self            # the reference's io.py reaches ffi as my_lib._my_lib.ffi
        self.ffi = self._ffi

    @property
    def _c(self):
        return _lib.load()

    # ---- /root/reference/src/my_lib.h:33-42
    def my_lib_ComputeChainObjfAndDeriv(self, den_graph_ptr, supervision_ptr, nnet_output, results,
                                        nnet_output_deriv, xent_output_deriv,
                                        l2_regularize, leaky_hmm_coefficient, xent_regularize):
        y, rows, cols, ystride = _matrix(nnet_output, "nnet_output")
        if results.is_cuda or results.dtype != torch.float32 or results.numel() < 3 or not results.is_contiguous():
            raise RuntimeError("results must be a contiguous CPU float tensor with 3 elements")
        d, drows, dcols, dstride = _matrix(nnet_output_deriv, "nnet_output_deriv")
        if (drows, dcols) != (rows, cols):
            raise RuntimeError("nnet_output_deriv must have the shape of nnet_output")
        x, xstride = None, 0
        if not _is_null(xent_output_deriv):
            # the reference's copy_to_mat resizes this tensor (src/my_lib_chain.cpp:22-29)
            if tuple(xent_output_deriv.shape) != (rows, cols) or not xent_output_deriv.is_contiguous():
                xent_output_deriv.resize_(rows, cols)
            x, _, _, xstride = _matrix(xent_output_deriv, "xent_output_deriv")
        dev = nnet_output.device.index
        stream = torch.cuda.current_stream(dev).cuda_stream
        ok = self._c.my_lib_ComputeChainObjfAndDeriv(
            _ptr(den_graph_ptr), _ptr(supervision_ptr), y, rows, cols, ystride, results.data_ptr(),
            d, dstride, x, xstride, float(l2_regularize), float(leaky_hmm_coefficient), float(xent_regularize),
            dev, ctypes.c_void_p(stream))
        if ok != 1:
            # the reference aborts inside Kaldi here; an exception is the closest Python can get
            raise RuntimeError("ComputeChainObjfAndDeriv failed: " + _lib.last_error())
        return 1

    # ---- src/my_lib.h:48
    def my_lib_set_kaldi_device(self, tensor):
        dev = tensor.device.index if isinstance(tensor, torch.Tensor) else int(tensor)
        if self._c.my_lib_set_kaldi_device(0 if dev is None else dev) != 1:
            raise RuntimeError(_lib.last_error())

    # ---- src/my_lib.h:29-30
    def my_lib_denominator_graph_new(self, rxfilename, num_pdf):
        if isinstance(rxfilename, ctypes.Array):
            rxfilename = rxfilename.value
        elif isinstance(rxfilename, str):
            rxfilename = rxfilename.encode()
        p = self._c.my_lib_denominator_graph_new(rxfilename, int(num_pdf))
        return p if p else self._ffi.NULL

    def my_lib_denominator_graph_free(self, ptr):
        if not _is_null(ptr):
            self._c.my_lib_denominator_graph_free(ptr)

    # ---- src/my_lib.h:19-25
    def my_lib_supervision_new(self, reader_ptr):
        p = self._c.my_lib_supervision_new(_ptr(reader_ptr))
        return p if p else self._ffi.NULL

    def my_lib_supervision_rand_new(self, reader_ptr):
        p = self._c.my_lib_supervision_rand_new(_ptr(reader_ptr))
        return p if p else self._ffi.NULL

    def my_lib_supervision_free(self, ptr):
        if not _is_null(ptr):
            self._c.my_lib_supervision_free(ptr)

    def my_lib_supervision_num_pdf(self, ptr):
        return self._c.my_lib_supervision_num_pdf(_ptr(ptr))

    def my_lib_supervision_num_sequence(self, ptr):
        return self._c.my_lib_supervision_num_sequence(_ptr(ptr))

    def my_lib_supervision_num_frame(self, ptr):
        return self._c.my_lib_supervision_num_frame(_ptr(ptr))

    def __getattr__(self, name):
        if name.startswith(("my_lib_example", "print_key_length", "my_lib_test_chain")):
            raise NotImplementedError(
                name + ": Kaldi egs readers / Kaldi's own test are outside the chain-loss hot path "
                       "(SURVEY.md section 8); build supervisions with torchain.io.Supervision.from_arrays "
                       "/ from_kaldi_bytes / from_chain_egs")
        raise AttributeError(name)


my_lib = _MyLib()
