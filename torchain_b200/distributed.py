"""Sequence-sharded multi-GPU use of the chain loss.

The minibatch shards by sequence: every recursion is independent per sequence,
the denominator graph is replicated, and only ``ChainResults.data`` =
``[objf, l2_term, weight]`` is summed over ranks.  This finishes what the
reference's ``example/chime5/parallel_train.py:55-75`` attempted with Python
threads (rows ``[loss, objf, l2, weight]`` gathered and summed).  One process
per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests); the
derivative never leaves its rank.

On GPUs the three scalars never visit the host before they are reduced: the
library leaves them in device memory (``results_dev`` of
``torchain_ComputeChainObjfAndDeriv_strided``), NCCL sums that tensor in place on
the same stream, and ONE device-to-host copy delivers the minibatch totals.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .functions import ChainResults, _ChainLoss, _tsp_strides, chain_loss  # noqa: F401


def shard_bounds(num_sequences, rank, world_size):
    """Sequences [lo, hi) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(num_sequences), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(x, num_sequences, rank, world_size):
    """Rows of a frame-major (T*B, P) matrix that belong to `rank`'s sequences,
    as a new contiguous (T*b, P) matrix (row = t*b + local sequence)."""
    lo, hi = shard_bounds(num_sequences, rank, world_size)
    T = x.shape[0] // num_sequences
    return x.view(T, num_sequences, -1)[:, lo:hi, :].reshape(T * (hi - lo), -1)


def _active(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def all_reduce_results(results, group=None):
    """Sum objf / l2_term / weight of a host-side ChainResults over the ranks of `group`, in place.
    (The GPU path below reduces on the device instead; this is for results that already are on the host.)"""
    if not _active(group):
        return results
    buf = results.data.to(torch.float64)
    if dist.get_backend(group) == "nccl":
        buf = buf.cuda()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    results.data.copy_(buf.cpu().to(torch.float32))
    return results


def sharded_step(den_graph, supervision, nnet_output, deriv, results_dev, l2_regularize, leaky_hmm_coefficient,
                 group=None, deriv_scale=-1.0, xent_deriv=None, xent_regularize=0.0):
    """One ComputeChainObjfAndDeriv on this rank's shard with the results left ON THE DEVICE, followed by
    the path's only exchange: ``results_dev`` (float32[3] CUDA tensor = objf, l2_term, weight) is summed
    over the ranks of `group` in place, on the current stream, without a host round trip.  Nothing
    synchronises; read ``results_dev`` (one .cpu()) when the minibatch totals are wanted on the host.
    `deriv` (and `xent_deriv`) receive this rank's derivative times `deriv_scale`."""
    lib = _lib.load()
    den_ptr = getattr(den_graph, "ptr", den_graph)
    sup_ptr = getattr(supervision, "ptr", supervision)
    n_seq = lib.my_lib_supervision_num_sequence(sup_ptr)
    if n_seq <= 0:
        raise RuntimeError("ComputeChainObjfAndDeriv failed: null or empty supervision")
    assert results_dev.is_cuda and results_dev.dtype == torch.float32 and results_dev.numel() >= 3
    if nnet_output.dim() == 2:
        n_frame, n_pdf = nnet_output.shape[0] // n_seq, nnet_output.shape[1]
    else:
        n_frame, n_pdf = nnet_output.shape[2], nnet_output.shape[1]
    dev = nnet_output.device.index
    stream = torch.cuda.current_stream(dev).cuda_stream
    ok = lib.torchain_ComputeChainObjfAndDeriv_strided(
        den_ptr, sup_ptr, nnet_output.data_ptr(), n_frame, n_seq, n_pdf, _tsp_strides(nnet_output, n_seq),
        None, results_dev.data_ptr(), deriv.data_ptr(), _tsp_strides(deriv, n_seq),
        xent_deriv.data_ptr() if xent_deriv is not None else None,
        _tsp_strides(xent_deriv, n_seq) if xent_deriv is not None else None,
        float(l2_regularize), float(leaky_hmm_coefficient), float(xent_regularize), float(deriv_scale),
        dev, ctypes.c_void_p(stream))
    if ok != 1:
        raise RuntimeError("ComputeChainObjfAndDeriv failed: " + _lib.last_error())
    if _active(group):
        dist.all_reduce(results_dev, op=dist.ReduceOp.SUM, group=group)
    return results_dev


class _ChainLossSharded(torch.autograd.Function):
    """`_ChainLoss` for one rank's shard: same gradient (rank-local, un-normalised, grad_output ignored, as the
    reference's backward, functions.py:106-115), results summed over the group on the device."""

    @staticmethod
    def forward(ctx, input, results, den_graph, supervision, l2_regularize, leaky_hmm_coefficient, group):
        assert input.is_cuda, "Only CUDA implementation is available"
        neg_grad = torch.empty_like(input)     # any (frame, sequence, pdf) strides are read / written in place
        res_dev = torch.empty(3, device=input.device, dtype=torch.float32)
        sharded_step(den_graph, supervision, input, neg_grad, res_dev, l2_regularize, leaky_hmm_coefficient, group,
                     -1.0)
        results.data.copy_(res_dev)            # ONE device-to-host copy: the minibatch totals
        ctx.neg_grad = neg_grad
        return input.new_tensor([float(results.loss)])

    @staticmethod
    def backward(ctx, grad_output):
        return ctx.neg_grad, None, None, None, None, None, None


def chain_loss_distributed(input, den_graph, supervision, group=None, l2_regularize=0.0,
                           leaky_hmm_coefficient=1e-5):
    """`chain_loss` on this rank's shard; the returned ChainResults (and the loss value) describe the whole
    minibatch.  Gradients stay rank-local, un-normalised, exactly as the reference's backward produces
    them.  (The xent branch is single-process only: use `chain_loss` and `all_reduce_results` for it.)"""
    results = ChainResults()
    loss = _ChainLossSharded.apply(input, results, den_graph, supervision, l2_regularize, leaky_hmm_coefficient, group)
    return loss, results
