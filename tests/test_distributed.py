"""N>1 path on CPU: world_size-2 gloo processes each own a sequence shard; the summed
ChainResults and the concatenated derivatives must equal the single-rank result on the
merged minibatch.  The per-shard arithmetic here is the ORACLE
(tests may use it as the checker); what is under test is the host-side sharding /
merging / reduction logic of torchain_b200.distributed and torchain_b200.synth."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from torchain_b200 import synth
from torchain_b200.distributed import all_reduce_results, shard_bounds, shard_rows
from torchain_b200.functions import ChainResults

HERE = os.path.dirname(os.path.abspath(__file__))


def _case():
    den = synth.make_den_fst(21, 60, 17, mean_deg=3.0)
    Bs, T = [3, 2], 6
    sups = [synth.make_supervision(30 + k, den, b, T, num_paths=3, boundary_final_costs=True)
            for k, b in enumerate(Bs)]
    ys = [synth.make_nnet_output(40 + k, b, T, den.num_pdfs) for k, b in enumerate(Bs)]
    return den, sups, ys, Bs, T


def _worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    den, sups, ys, Bs, T = _case()
    r = ol.chain(ol.OracleDen(den), ol.OracleSup(sups[rank]), ys[rank], l2=0.01, leaky=0.1, mode="fp64")
    res = ChainResults()
    res.data[0], res.data[1], res.data[2] = r["objf"], r["l2_term"], r["weight"]
    all_reduce_results(res)
    out[rank] = (res.data.numpy().copy(), r["deriv"])
    dist.destroy_process_group()


def test_two_rank_shards_equal_merged_minibatch():
    den, sups, ys, Bs, T = _case()
    merged = synth.append_supervisions(sups)
    y_all = synth.interleave_rows(ys, Bs)
    full = ol.chain(ol.OracleDen(den), ol.OracleSup(merged), y_all, l2=0.01, leaky=0.1, mode="fp64")
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in range(2):
        data, _ = out[rank]
        assert data[0] == pytest.approx(full["objf"], rel=1e-6)
        assert data[1] == pytest.approx(full["l2_term"], rel=1e-6)
        assert data[2] == pytest.approx(full["weight"])
    d_all = synth.interleave_rows([out[0][1], out[1][1]], Bs)
    assert np.abs(d_all - full["deriv"]).max() < 1e-9


def test_shard_bounds_and_rows():
    assert [shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    B, T, P = 5, 3, 2
    x = torch.arange(T * B * P, dtype=torch.float32).view(T * B, P)
    parts = [shard_rows(x, B, r, 2) for r in range(2)]
    assert parts[0].shape == (T * 3, P) and parts[1].shape == (T * 2, P)
    back = synth.interleave_rows([p.numpy() for p in parts], [3, 2])
    assert np.array_equal(back, x.numpy())


def test_append_supervisions_matches_direct_generation_semantics():
    """Merging per-sequence supervisions gives a valid merged FST whose numerator log-prob is
    the sum of the parts' (Concat multiplies path weights)."""
    den, sups, ys, Bs, T = _case()
    oden = ol.OracleDen(den)
    merged = synth.append_supervisions(sups)
    y_all = synth.interleave_rows(ys, Bs)
    tot = ol.chain(oden, ol.OracleSup(merged), y_all, leaky=0.1, mode="fp64", want_deriv=False)["num"]
    parts = sum(ol.chain(oden, ol.OracleSup(sp), y, leaky=0.1, mode="fp64", want_deriv=False)["num"]
                for sp, y in zip(sups, ys))
    assert tot == pytest.approx(parts, rel=1e-7)   # merged arc costs are float32 sums final(f) + cost


@pytest.mark.gpu
def test_gpu_shards_equal_full_batch():
    """Same identity on the CUDA path (single GPU, the two shards run one after the other)."""
    import ctypes
    from torchain_b200 import _lib, io
    den, sups, ys, Bs, T = _case()
    merged = synth.append_supervisions(sups)
    y_all = synth.interleave_rows(ys, Bs)
    lib = _lib.load()
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)

    def run(sup, y):
        s = io.Supervision.from_arrays(sup.weight, sup.num_sequences, sup.frames_per_sequence, sup.label_dim,
                                       sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)
        yt = torch.from_numpy(y).cuda()
        d = torch.empty_like(yt)
        res = np.zeros(3, dtype=np.float32)
        ok = lib.my_lib_ComputeChainObjfAndDeriv(g.ptr, s.ptr, yt.data_ptr(), y.shape[0], y.shape[1], y.shape[1],
                                                 res.ctypes.data_as(ctypes.c_void_p), d.data_ptr(), y.shape[1],
                                                 None, 0, 0.01, 0.1, 0.0, 0, None)
        assert ok == 1, _lib.last_error()
        torch.cuda.synchronize()
        return res, d.cpu().numpy()

    full, dfull = run(merged, y_all)
    parts = [run(sp, y) for sp, y in zip(sups, ys)]
    assert parts[0][0][0] + parts[1][0][0] == pytest.approx(full[0], rel=1e-5)
    assert parts[0][0][2] + parts[1][0][2] == pytest.approx(full[2])
    d_all = synth.interleave_rows([parts[0][1], parts[1][1]], Bs)
    assert np.abs(d_all - dfull).max() < 1e-4 * np.abs(dfull).max()


def _nccl_worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from torchain_b200 import io
    from torchain_b200.distributed import chain_loss_distributed
    den, sups, ys, Bs, T = _case()
    sp = sups[rank]
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    s = io.Supervision.from_arrays(sp.weight, sp.num_sequences, sp.frames_per_sequence, sp.label_dim, sp.arc_begin,
                                   sp.arc_dst, sp.arc_label, sp.arc_cost, sp.final_cost)
    y = torch.from_numpy(ys[rank]).cuda().requires_grad_(True)
    loss, results = chain_loss_distributed(y, g, s, l2_regularize=0.01, leaky_hmm_coefficient=0.1)
    loss.backward()
    torch.cuda.synchronize()
    out[rank] = (results.data.numpy().copy(), float(loss), y.grad.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_chain_loss_distributed_two_gpus():
    """The distributed op itself under NCCL (run with `gpurun --gpus 2`): every rank's ChainResults describe the
    merged minibatch (objf / l2_term / weight summed on the device), gradients stay rank-local."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    den, sups, ys, Bs, T = _case()
    merged = synth.append_supervisions(sups)
    y_all = synth.interleave_rows(ys, Bs)
    full = ol.chain(ol.OracleDen(den), ol.OracleSup(merged), y_all, l2=0.01, leaky=0.1, mode="fp64")
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in range(2):
        data, loss, _ = out[rank]
        assert data[0] == pytest.approx(full["objf"], rel=1e-5)
        assert data[1] == pytest.approx(full["l2_term"], rel=1e-5)
        assert data[2] == pytest.approx(full["weight"])
        assert loss == pytest.approx(-full["objf"] / full["weight"], rel=1e-5)
    d_all = synth.interleave_rows([out[0][2], out[1][2]], Bs)
    assert np.abs(d_all + full["deriv"]).max() < 1e-4 * np.abs(full["deriv"]).max()


@pytest.mark.gpu
def test_gpu_sharded_step_leaves_results_on_the_device():
    """Single process: the sharded op is the plain op plus a no-op reduction; results arrive in device memory and
    the call never synchronises."""
    from torchain_b200 import io
    from torchain_b200.distributed import chain_loss_distributed, sharded_step
    den, sups, ys, Bs, T = _case()
    sp, y_np = sups[0], ys[0]
    ref = ol.chain(ol.OracleDen(den), ol.OracleSup(sp), y_np, l2=0.01, leaky=0.1, mode="fp64")
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    s = io.Supervision.from_arrays(sp.weight, sp.num_sequences, sp.frames_per_sequence, sp.label_dim, sp.arc_begin,
                                   sp.arc_dst, sp.arc_label, sp.arc_cost, sp.final_cost)
    y = torch.from_numpy(y_np).cuda()
    d = torch.empty_like(y)
    res = torch.zeros(3, device="cuda")
    sharded_step(g, s, y, d, res, 0.01, 0.1)
    assert res.cpu()[0].item() == pytest.approx(ref["objf"], rel=1e-5)
    assert np.abs(d.cpu().numpy() + ref["deriv"]).max() < 1e-4 * np.abs(ref["deriv"]).max()
    yg = y.clone().requires_grad_(True)
    loss, results = chain_loss_distributed(yg, g, s, l2_regularize=0.01, leaky_hmm_coefficient=0.1)
    (loss * 7.0).backward()
    assert float(results.data[0]) == pytest.approx(ref["objf"], rel=1e-5)
    assert np.abs(yg.grad.cpu().numpy() + ref["deriv"]).max() < 1e-4 * np.abs(ref["deriv"]).max()
