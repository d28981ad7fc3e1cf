"""LF-MMI (chain) loss as a PyTorch autograd Function.

Same public surface and behaviour as the reference's ``torchain/functions.py``
(/root/reference/torchain/functions.py:9-138): ``ChainResults``, ``to2d``,
``chain_loss(input, den_graph, supervision, l2_regularize=0.0,
leaky_hmm_coefficient=1e-5, xent_regularize=0.0, xent_input=None,
kaldi_way=False) -> (loss, results)``.  Underneath, one call into
libtorchain_b200.so replaces ``my_lib.my_lib_ComputeChainObjfAndDeriv``
(functions.py:82-86 -> src/my_lib_chain.cpp:104-136).

Quirks kept on purpose: backward ignores ``grad_output``
and does not divide by ``weight``; the returned scalar is ``-objf/weight``
without the l2 term; ``kaldi_way=False`` runs the loss a second time on
``xent_input`` and overwrites both ``results`` and the LF-MMI gradient.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from . import io


class ChainResults:
    """Reference: functions.py:9-19.  ``data`` = CPU float32[3] = objf, l2_term, weight."""

    def __init__(self):
        self.data = torch.zeros(3)

    def __repr__(self):
        return "ChainResults(loss=%f, objf=%f, l2_term=%f, weight=%lf)" % (
            self.loss, self.data[0], self.data[1], self.data[2])

    @property
    def loss(self):
        return -self.data[0] / self.data[2]


def _tsp_strides(x, n_seq):
    """Element strides of `x` over (frame, sequence, pdf) as a ctypes int64[3].

    2-D (T*B, P), row = t*B + b (what the reference's to2d produces): (B*stride0, stride0, stride1).
    3-D (B, P, T), the shape the reference's models emit (functions.py:118-125): (stride2, stride0, stride1).
    Any strides are accepted: the kernels address elements through these three numbers."""
    if x.dim() == 2:
        st = (n_seq * x.stride(0), x.stride(0), x.stride(1))
    else:
        st = (x.stride(2), x.stride(0), x.stride(1))
    return (ctypes.c_int64 * 3)(*st)


def _compute(den_graph, supervision, nnet_output, results, deriv, xent_deriv,
             l2_regularize, leaky_hmm_coefficient, xent_regularize, deriv_scale):
    """One ComputeChainObjfAndDeriv on the tensors' device / current stream.

    `nnet_output` (and the derivative buffers, which have its shape) is either the reference's
    (T*B, P) matrix or a (B, P, T) tensor read in place -- no to2d copy."""
    lib = _lib.load()
    if nnet_output.dim() not in (2, 3):
        raise RuntimeError("dim != 2")  # reference: src/common.hpp:110
    if nnet_output.dtype != torch.float32:
        raise RuntimeError("nnet_output must be float32")
    n_seq = lib.my_lib_supervision_num_sequence(supervision)
    if n_seq <= 0:
        raise RuntimeError("ComputeChainObjfAndDeriv failed: null or empty supervision")
    if nnet_output.dim() == 2:
        if nnet_output.shape[0] % n_seq != 0:
            raise RuntimeError("ComputeChainObjfAndDeriv failed: nnet_output rows != "
                               "num_sequences * frames_per_sequence")
        n_frame, n_pdf = nnet_output.shape[0] // n_seq, nnet_output.shape[1]
    else:
        if nnet_output.shape[0] != n_seq:
            raise RuntimeError("ComputeChainObjfAndDeriv failed: nnet_output batch != num_sequences")
        n_frame, n_pdf = nnet_output.shape[2], nnet_output.shape[1]
    dev = nnet_output.device.index
    if dev is None:
        dev = torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev).cuda_stream
    res = results.data
    assert res.dtype == torch.float32 and res.numel() >= 3 and not res.is_cuda and res.is_contiguous()
    def call():
        return lib.torchain_ComputeChainObjfAndDeriv_strided(
            den_graph, supervision,
            nnet_output.data_ptr(), n_frame, n_seq, n_pdf, _tsp_strides(nnet_output, n_seq),
            res.data_ptr(), None,
            deriv.data_ptr(), _tsp_strides(deriv, n_seq),
            xent_deriv.data_ptr() if xent_deriv is not None else None,
            _tsp_strides(xent_deriv, n_seq) if xent_deriv is not None else None,
            float(l2_regularize), float(leaky_hmm_coefficient), float(xent_regularize), float(deriv_scale),
            dev, ctypes.c_void_p(stream))

    ok = call()
    if ok != 1 and "out of memory" in _lib.last_error():
        # the library's workspace (the alpha store) is allocated with cudaMalloc, which cannot see the blocks
        # torch's caching allocator holds free: hand them back and try once more
        torch.cuda.empty_cache()
        ok = call()
    if ok != 1:
        raise RuntimeError("ComputeChainObjfAndDeriv failed: " + _lib.last_error())


def _stream_args(x):
    dev = x.device.index
    if dev is None:
        dev = torch.cuda.current_device()
    return dev, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _relayout_ok(x):
    """A contiguous (B, P, T) tensor the tiled relayout kernels accept (layout_kernel.cu)."""
    return (x.dim() == 3 and x.is_contiguous() and x.shape[2] > 1 and x.shape[2] <= 1599 and x.shape[0] <= 65535
            and x.dtype == torch.float32 and x.numel() > 0)


def _rows_of(x):
    """to2d by the library's tiled kernel: (B, P, T) contiguous -> (T*B, P), row = t*B + b."""
    B, P, T = x.shape
    rows = torch.empty(T * B, P, dtype=x.dtype, device=x.device)
    dev, stream = _stream_args(x)
    if _lib.load().torchain_to2d(x.data_ptr(), B, P, T, rows.data_ptr(), P, dev, stream) != 1:
        raise RuntimeError("to2d failed: " + _lib.last_error())
    return rows


def _rows_to(rows, like):
    """A (T*B, P) derivative brought to the shape / layout of `like` (what autograd's backward of to2d does)."""
    if like.dim() == 2:
        return rows
    B, P, T = like.shape
    if _relayout_ok(like):
        out = torch.empty_like(like)
        dev, stream = _stream_args(rows)
        if _lib.load().torchain_from2d(rows.data_ptr(), rows.stride(0), B, P, T, out.data_ptr(), 1.0, dev, stream) != 1:
            raise RuntimeError("from2d failed: " + _lib.last_error())
        return out
    out = torch.empty_like(like)
    out.copy_(rows.view(T, B, P).permute(1, 2, 0))
    return out


def _evaluate(den_graph, supervision, x, results, like_mmi, like_xent,
              l2_regularize, leaky_hmm_coefficient, xent_regularize):
    """One loss evaluation on `x`; returns the NEGATED derivatives in the shape and layout of
    `like_mmi` / `like_xent` (None without xent).

    Where the reference always goes through to2d (a permute+contiguous copy in, a permuted gradient
    copy out), a 3-D (B, P, T) network output takes one of two cheaper routes:
      * pdf stride 1 (e.g. a (B, T, P) tensor viewed as (B, P, T)), or any other strided view: the
        kernels read it and write the gradient in place through (frame, sequence, pdf) strides;
      * contiguous (B, P, T), frames innermost: single elements 4*T bytes apart would cost the
        recurrence 32 sectors per load, so the matrix goes through the library's tiled relayout
        kernels (HBM speed) on the way in and the gradient on the way out."""
    want_xent = like_xent is not None
    if _relayout_ok(x):
        rows = _rows_of(x)
        d2 = torch.empty_like(rows)
        x2 = torch.empty_like(rows) if want_xent else None
        _compute(den_graph, supervision, rows, results, d2, x2,
                 l2_regularize, leaky_hmm_coefficient, xent_regularize, -1.0)
        return _rows_to(d2, like_mmi), (_rows_to(x2, like_xent) if want_xent else None)
    d = torch.empty_like(like_mmi)
    xd = torch.empty_like(like_xent) if want_xent else None
    _compute(den_graph, supervision, x, results, d, xd,
             l2_regularize, leaky_hmm_coefficient, xent_regularize, -1.0)
    return d, xd


class _ChainLoss(Function):
    """Lattice-free MMI loss (reference: functions.py:22-115).

    input        (num_sequences*frames_per_sequence, num_pdfs) CUDA float tensor, rows
                 ordered all sequences of frame 0, all sequences of frame 1, ...; or the
                 network output as it is, (num_sequences, num_pdfs, frames_per_sequence) --
                 see _evaluate: read in place, or moved by the library's own relayout kernels,
                 instead of to2d's permute+contiguous and the permute in its autograd backward
                 (functions.py:118-125)
    den_graph    io.DenominatorGraph (or its raw pointer)
    supervision  io.Supervision (or its raw pointer)
    results      ChainResults, filled with objf / l2_term / weight
    """

    @staticmethod
    def forward(ctx, input, xent_input, results, den_graph, supervision,
                l2_regularize, leaky_hmm_coefficient, xent_regularize=0.0, kaldi_way=False):
        assert input.is_cuda, "Only CUDA implementation is available"
        if isinstance(supervision, io.Supervision):
            supervision = supervision.ptr
        if isinstance(den_graph, io.DenominatorGraph):
            den_graph = den_graph.ptr

        # The library writes the derivative already negated (deriv_scale = -1), so
        # backward() hands it out without launching a negate kernel.
        use_xent = xent_input is not None and xent_regularize != 0.0
        like_xent = xent_input if use_xent else None
        neg_mmi_grad, neg_xent_grad = _evaluate(den_graph, supervision, input, results, input, like_xent,
                                                l2_regularize, leaky_hmm_coefficient, xent_regularize)
        if use_xent and not kaldi_way:
            # reference functions.py:96-103: second full call on the xent output,
            # overwriting results and the LF-MMI gradient buffer
            neg_mmi_grad, neg_xent_grad = _evaluate(den_graph, supervision, xent_input, results, input, like_xent,
                                                    l2_regularize, leaky_hmm_coefficient, xent_regularize)
        ctx.neg_mmi_grad = neg_mmi_grad
        if use_xent:
            ctx.neg_xent_grad = neg_xent_grad.mul_(xent_regularize)
        return input.new_tensor([float(results.loss)])

    @staticmethod
    def backward(ctx, grad_output):
        neg_xent_grad = getattr(ctx, "neg_xent_grad", None)
        return (ctx.neg_mmi_grad, neg_xent_grad,
                None, None, None,
                None, None, None,
                None)


class _To2d(Function):
    """to2d of a contiguous CUDA (B, C, T) tensor through the library's relayout kernels."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        ctx.device = x.device
        return _rows_of(x)

    @staticmethod
    def backward(ctx, grad_rows):
        B, P, T = ctx.shape
        # any gradient layout is accepted, as by the backward of the reference's permute().contiguous().view():
        # a row-broadcast gradient (to2d(x).sum(0)) arrives expanded with strides (0, 1), a double one from .double()
        if (grad_rows.stride(1) != 1 or grad_rows.stride(0) < P or grad_rows.dtype != torch.float32
                or not grad_rows.is_cuda):
            grad_rows = grad_rows.to(device=ctx.device, dtype=torch.float32).contiguous()
        out = grad_rows.new_empty(B, P, T)
        dev, stream = _stream_args(grad_rows)
        if _lib.load().torchain_from2d(grad_rows.data_ptr(), grad_rows.stride(0), B, P, T, out.data_ptr(), 1.0,
                                       dev, stream) != 1:
            raise RuntimeError("from2d failed: " + _lib.last_error())
        return out


def to2d(x):
    """(B, C, T) -> (T*B, C), row = t*B + b (reference: functions.py:118-125).

    Same values as the reference's ``x.permute(2, 0, 1).contiguous().view(-1, C)``; a contiguous
    CUDA float tensor is moved by the tiled kernels of layout_kernel.cu (and so is its gradient)."""
    if x.dim() == 3:
        if x.is_cuda and _relayout_ok(x):
            return _To2d.apply(x)
        n_pdf = x.shape[1]
        x = x.permute(2, 0, 1).contiguous().view(-1, n_pdf)
    assert x.dim() == 2
    return x


def chain_loss(input, den_graph, supervision,
               l2_regularize=0.0, leaky_hmm_coefficient=1e-5,
               xent_regularize=0.0, xent_input=None, kaldi_way=False):
    """Reference: functions.py:128-138.  Returns ``(loss, ChainResults)``.

    The reference copies a (B, C, T) input to (T*B, C) with ``to2d`` first; here both shapes go
    to the autograd op as they are (see _evaluate), the values are the same."""
    assert input.dim() in (2, 3)
    if xent_input is not None:
        assert xent_input.dim() in (2, 3)
    results = ChainResults()
    loss = _ChainLoss.apply(input, xent_input, results, den_graph, supervision,
                            l2_regularize, leaky_hmm_coefficient, xent_regularize, kaldi_way)
    return loss, results
