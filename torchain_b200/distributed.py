"""Sequence-sharded multi-GPU use of the chain loss.

The minibatch shards by sequence: every recursion is independent per sequence,
the denominator graph is replicated, and only ``ChainResults.data`` =
``[objf, l2_term, weight]`` is summed over ranks.  This finishes what the
reference's ``example/chime5/parallel_train.py:55-75`` attempted with Python
threads (rows ``[loss, objf, l2, weight]`` gathered and summed).  One process
per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests); the
derivative never leaves its rank.
"""
import torch
import torch.distributed as dist

from .functions import ChainResults, chain_loss


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


def all_reduce_results(results, group=None):
    """Sum objf / l2_term / weight over the ranks of `group`, in place."""
    if not (dist.is_available() and dist.is_initialized()):
        return results
    backend = dist.get_backend(group)
    buf = results.data.to(torch.float64)
    if backend == "nccl":
        buf = buf.cuda()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    results.data.copy_(buf.cpu().to(torch.float32))
    return results


def chain_loss_distributed(input, den_graph, supervision, group=None, **kwargs):
    """`chain_loss` on this rank's shard; the returned ChainResults (and the loss
    value) describe the whole minibatch.  Gradients stay rank-local, un-normalised,
    exactly as the reference's backward produces them."""
    loss, results = chain_loss(input, den_graph, supervision, **kwargs)
    all_reduce_results(results, group)
    return loss.new_tensor([float(results.loss)]) + (loss - loss.detach()), results
