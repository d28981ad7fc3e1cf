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


def _compute(den_graph, supervision, nnet_output, results, deriv, xent_deriv,
             l2_regularize, leaky_hmm_coefficient, xent_regularize, deriv_scale):
    """One ComputeChainObjfAndDeriv on the tensors' device / current stream."""
    lib = _lib.load()
    if nnet_output.dim() != 2:
        raise RuntimeError("dim != 2")  # reference: src/common.hpp:110
    if nnet_output.dtype != torch.float32 or nnet_output.stride(1) != 1:
        raise RuntimeError("nnet_output must be float32 with unit column stride")
    dev = nnet_output.device.index
    if dev is None:
        dev = torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev).cuda_stream
    res = results.data
    assert res.dtype == torch.float32 and res.numel() >= 3 and not res.is_cuda and res.is_contiguous()
    ok = lib.torchain_ComputeChainObjfAndDeriv_ex(
        den_graph, supervision,
        nnet_output.data_ptr(), nnet_output.shape[0], nnet_output.shape[1], nnet_output.stride(0),
        res.data_ptr(), None,
        deriv.data_ptr(), deriv.stride(0),
        xent_deriv.data_ptr() if xent_deriv is not None else None,
        xent_deriv.stride(0) if xent_deriv is not None else 0,
        float(l2_regularize), float(leaky_hmm_coefficient), float(xent_regularize), float(deriv_scale),
        dev, ctypes.c_void_p(stream))
    if ok != 1:
        raise RuntimeError("ComputeChainObjfAndDeriv failed: " + _lib.last_error())


class _ChainLoss(Function):
    """Lattice-free MMI loss (reference: functions.py:22-115).

    input        (num_sequences*frames_per_sequence, num_pdfs) CUDA float tensor, rows
                 ordered all sequences of frame 0, all sequences of frame 1, ...
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
        neg_mmi_grad = torch.empty_like(input)
        use_xent = xent_input is not None and xent_regularize != 0.0
        neg_xent_grad = torch.empty_like(xent_input) if use_xent else None

        _compute(den_graph, supervision, input, results, neg_mmi_grad, neg_xent_grad,
                 l2_regularize, leaky_hmm_coefficient, xent_regularize, -1.0)
        ctx.neg_mmi_grad = neg_mmi_grad
        if use_xent:
            if not kaldi_way:
                # reference functions.py:96-103: second full call on the xent output,
                # overwriting results and the LF-MMI gradient buffer
                _compute(den_graph, supervision, xent_input, results, neg_mmi_grad, neg_xent_grad,
                         l2_regularize, leaky_hmm_coefficient, xent_regularize, -1.0)
            ctx.neg_xent_grad = neg_xent_grad.mul_(xent_regularize)
        return input.new_tensor([float(results.loss)])

    @staticmethod
    def backward(ctx, grad_output):
        neg_xent_grad = getattr(ctx, "neg_xent_grad", None)
        return (ctx.neg_mmi_grad, neg_xent_grad,
                None, None, None,
                None, None, None,
                None)


def to2d(x):
    """(B, C, T) -> (T*B, C), row = t*B + b (reference: functions.py:118-125)."""
    if x.dim() == 3:
        n_pdf = x.shape[1]
        x = x.permute(2, 0, 1).contiguous().view(-1, n_pdf)
    assert x.dim() == 2
    return x


def chain_loss(input, den_graph, supervision,
               l2_regularize=0.0, leaky_hmm_coefficient=1e-5,
               xent_regularize=0.0, xent_input=None, kaldi_way=False):
    """Reference: functions.py:128-138.  Returns ``(loss, ChainResults)``."""
    input = to2d(input)
    if xent_input is not None:
        xent_input = to2d(xent_input)
    results = ChainResults()
    loss = _ChainLoss.apply(input, xent_input, results, den_graph, supervision,
                            l2_regularize, leaky_hmm_coefficient, xent_regularize, kaldi_way)
    return loss, results
