#!/usr/bin/env python
"""bench.py -- chain loss forward+backward throughput (frames/s) on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config chime5]

Own arm: one step = one ComputeChainObjfAndDeriv-equivalent call (numerator
forward-backward + denominator alpha/beta + derivative, incl. the negation the
reference's backward() does) over one synthetic minibatch of BASELINE.json
config 2 ("chime5": P=2928, N=10000, A~300k, B=128 x T=50 per GPU, leaky 0.1,
l2 5e-5).  N>1: one process per GPU (torchrun), each rank owns B sequences
(weak scaling), the 3 result scalars are all-reduced over NCCL every step by the
package's distributed op (torchain_b200.distributed.sharded_step).  The line also
carries: a parity block (the step's results against the fp64 oracle, every rank),
the step with a NEW Supervision handle per minibatch (fresh_supervision_*), and
the other BASELINE.json configurations on the same launch geometry as flat keys
(cfg3_strong_*: 512 sequences split over the ranks; cfg4_*: large tree).

Reference arm (--impl reference): the Kaldi-CPU-equivalent restatement
(oracle/, kaldi_fp32 mode) on the host cores, bounded sample, rank 0 only.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chain_loss_fwd_bwd_frames_per_sec"
UNIT = "frames/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def host_cores():
    """Host threads usable by the CPU baseline (torchrun pins OMP_NUM_THREADS=1, so ask the OS)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch pair of the denominator kernels from the
    committed `ncu --set full` capture (profiles/ncu_den_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_den_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def make_workload(name, B, T, rank):
    from torchain_b200 import synth
    c = dict(synth.CONFIGS[name])
    if B:
        c["B"] = B
    if T:
        c["T"] = T
    den = (synth.make_den_fst_tied if c.get("tied") else synth.make_den_fst)(1234, c["N"], c["P"],
                                                                              mean_deg=c["mean_deg"])
    if c.get("hubs"):
        den = synth.add_hubs(den, c["hubs"])
    sup = synth.make_supervision(1235 + 1000 * rank, den, c["B"], c["T"], num_paths=4, weight=1.0)
    c.update(den=den, sup=sup)
    return c


def alg_bytes(c):
    """algorithmic bytes per sequence-frame."""
    hbm = 4 * (3 * c["P"] + 2 * (c["N"] + 1))
    gather = 16 * c["den"].num_arcs
    return hbm, gather


def cpu_baseline(c, seconds_budget=20.0, threads=None):
    """Oracle (kaldi_fp32) on a bounded sample: the same den graph and options, the first
    Bs sequences of a T-frame minibatch.  Returns (frames/s, cores, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from torchain_b200 import synth
    den = c["den"]
    ncores = threads or host_cores()
    oden = ol.OracleDen(den)
    T = c["T"]
    Bs = min(c["B"], max(8, 2 * ncores))
    best = None
    while True:
        sup = synth.make_supervision(77, den, Bs, T, num_paths=4, weight=1.0)
        y = synth.make_nnet_output(78, Bs, T, c["P"], scale=2.0)
        osup = ol.OracleSup(sup)
        t0 = time.perf_counter()
        ol.chain(oden, osup, y, l2=c["l2"], leaky=c["leaky"], mode="kaldi_fp32", threads=ncores)
        dt = time.perf_counter() - t0
        best = (Bs * T / dt, Bs, dt)
        if dt * 4 > seconds_budget or Bs >= c["B"]:
            break
        Bs = min(c["B"], Bs * 2)
    fps, Bs, dt = best
    return fps, ncores, "B=%d of %d sequences x T=%d frames, same den graph/options, %.1f s" % (Bs, c["B"], T, dt)


def parity_check(c, sup, y_np, res, neg_deriv_np, threads):
    """The step's results against the fp64 oracle on the same inputs (outside every timed region).
    Tolerances are BASELINE.json's: objf, l2_term 1e-5 relative; derivative 1e-4 of the largest magnitude."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    t0 = time.perf_counter()
    r = ol.chain(ol.OracleDen(c["den"]), ol.OracleSup(sup), y_np, l2=c["l2"], leaky=c["leaky"], mode="fp64",
                 threads=threads)
    e_objf = abs(float(res[0]) - r["objf"]) / max(abs(r["objf"]), 1e-30)
    e_l2 = abs(float(res[1]) - r["l2_term"]) / max(abs(r["l2_term"]), 1e-30)
    e_w = abs(float(res[2]) - r["weight"]) / max(abs(r["weight"]), 1e-30)
    e_d = float(np.abs(-neg_deriv_np.astype(np.float64) - r["deriv"]).max() / max(np.abs(r["deriv"]).max(), 1e-30))
    return {"objf_rel_err": e_objf, "l2_rel_err": e_l2, "weight_rel_err": e_w, "deriv_rel_err": e_d,
            "oracle_objf": r["objf"], "oracle_ok": bool(r["ok"]), "oracle_s": time.perf_counter() - t0,
            "ok": bool(r["ok"] and e_objf < 1e-5 and e_l2 < 1e-5 and e_w < 1e-6 and e_d < 1e-4)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = make_workload(args.config, args.batch, args.frames, 0)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from torchain_b200 import synth
    ncores = host_cores()
    oden = ol.OracleDen(c["den"])
    Bs = min(c["B"], max(8, 2 * ncores))
    T = c["T"]
    sup = synth.make_supervision(77, c["den"], Bs, T, num_paths=4, weight=1.0)
    y = synth.make_nnet_output(78, Bs, T, c["P"], scale=2.0)
    osup = ol.OracleSup(sup)
    for _ in range(args.warmup):
        ol.chain(oden, osup, y, l2=c["l2"], leaky=c["leaky"], mode="kaldi_fp32", threads=ncores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ol.chain(oden, osup, y, l2=c["l2"], leaky=c["leaky"], mode="kaldi_fp32", threads=ncores)
    dt = (time.perf_counter() - t0) / args.steps
    fps = Bs * T / dt
    sample = "B=%d of %d sequences x T=%d frames per step, same den graph/options" % (Bs, c["B"], T)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "chime5-shaped chain loss: P=%d N=%d A=%d B=%d T=%d leaky=%g l2=%g" % (
            c["P"], c["N"], c["den"].num_arcs, c["B"], c["T"], c["leaky"], c["l2"]),
            "note": "Kaldi-CPU-equivalent restatement (oracle kaldi_fp32); Kaldi itself is not buildable here"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


class Runner:
    """One workload resident on this rank's GPU: handles, nnet_output, derivative, and the step the bench times.

    A step is one ComputeChainObjfAndDeriv-equivalent call (derivative negated in-kernel, what the reference's
    backward() returns).  N=1: `torchain_ComputeChainObjfAndDeriv_ex` with the three results delivered to the host
    (the reference's call semantics: ChainResults.data is a CPU tensor).  N>1: the package's distributed op
    (`torchain_b200.distributed.sharded_step`): results stay on the device, NCCL sums them in place, one
    device-to-host read."""

    def __init__(self, torch, lib, c, dev, local, world, rank):
        from torchain_b200 import io
        self.torch, self.lib, self.c, self.dev, self.local, self.world = torch, lib, c, dev, local, world
        den, sup = c["den"], c["sup"]
        self.B, self.T, self.P = c["B"], c["T"], c["P"]
        self.g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                                 den.final_cost, den.num_pdfs)
        self.s = self.new_supervision(sup)
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        self.y = torch.randn(self.T * self.B, self.P, device=dev, generator=gen) * 2.0
        self.deriv = torch.empty_like(self.y)
        self.res_dev = torch.zeros(3, device=dev)
        self.res_host = np.zeros(3, dtype=np.float32)
        self.stream = torch.cuda.current_stream(dev)

    def new_supervision(self, sup):
        from torchain_b200 import io
        return io.Supervision.from_arrays(sup.weight, sup.num_sequences, sup.frames_per_sequence, sup.label_dim,
                                          sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)

    def step(self, host_results=False, s=None):
        # Multi-GPU steps take the results on the device (they are all-reduced first); the warm-up
        # steps ask for host results on every rank too, because only a synchronising call lets the
        # library read back its per-warp busy counters and calibrate the partition --
        # what the Python op (chain_loss) does on every call.
        c, s = self.c, (s or self.s)
        if self.world == 1 or host_results:
            ok = self.lib.torchain_ComputeChainObjfAndDeriv_ex(
                self.g.ptr, s.ptr, self.y.data_ptr(), self.T * self.B, self.P, self.y.stride(0),
                self.res_host.ctypes.data_as(ctypes.c_void_p), self.res_dev.data_ptr(),
                self.deriv.data_ptr(), self.deriv.stride(0), None, 0,
                c["l2"], c["leaky"], 0.0, -1.0, self.local, ctypes.c_void_p(self.stream.cuda_stream))
            if ok != 1:
                from torchain_b200 import _lib
                raise RuntimeError(_lib.last_error())
        else:
            from torchain_b200.distributed import sharded_step
            sharded_step(self.g, s, self.y, self.deriv, self.res_dev, c["l2"], c["leaky"], None, -1.0)
            self.res_host[:] = self.res_dev.cpu().numpy()   # the minibatch totals, one device-to-host read

    def sync_all(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, steps, warmup, step=None):
        """ms per step: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        step = step or self.step
        for _ in range(warmup):
            self.step(host_results=True)
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            step()
        e1.record(self.stream)
        self.sync_all()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps


def extra_config(torch, lib, name, B, T, dev, local, world, rank, steps=10):
    """(ms per step, whole-job frames/s, gather frac of the 32 MB-table L2 peak is left to the caller) of another
    BASELINE.json configuration on this launch geometry."""
    c = make_workload(name, B, T, rank)
    r = Runner(torch, lib, c, dev, local, world, rank)
    ms = r.time_steps(steps, 4)
    out = {"ms_per_step": ms, "frames_per_sec": world * c["B"] * c["T"] / (ms * 1e-3),
           "gather_gbs": alg_bytes(c)[1] * c["B"] * c["T"] / (ms * 1e-3) / 1e9,
           "objf": float(r.res_host[0]), "weight": float(r.res_host[2])}
    del r
    lib.torchain_supervision_pool_trim()
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor
    from torchain_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    c = make_workload(args.config, args.batch, args.frames, rank)
    den, sup = c["den"], c["sup"]
    B, T, P = c["B"], c["T"], c["P"]
    R = Runner(torch, lib, c, dev, local, world, rank)
    g, s, y, deriv, res_host, stream = R.g, R.s, R.y, R.deriv, R.res_host, R.stream
    step, sync_all = R.step, R.sync_all

    for _ in range(max(args.warmup, 3)):
        step(host_results=True)
    sync_all()
    lib.torchain_profile_enable(1)
    dummy_ms, dummy_n = ctypes.c_double(0), ctypes.c_int64(0)
    lib.torchain_profile_read(g.ptr, ctypes.byref(dummy_ms), ctypes.byref(dummy_n), 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.torchain_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = lib.torchain_kernel_launch_count() - launches0
    den_ms, den_n = ctypes.c_double(0), ctypes.c_int64(0)
    lib.torchain_profile_read(g.ptr, ctypes.byref(den_ms), ctypes.byref(den_n), 1)
    lib.torchain_profile_enable(0)
    ms_per_step = R.max_over_ranks(ms) / args.steps
    value = world * B * T / (ms_per_step * 1e-3)
    step_results = res_host.copy()

    # ---- the same step with the per-minibatch host work inside the timed region: a NEW Supervision handle every
    # step (validation, state times, per-sequence split, CSR both ways, device upload), as a training loop has it
    # (/root/reference/example/chime5/train_faster.py:116-141 takes a new supervision from its reader every step).
    # A worker thread builds and prefetches handle k+1 (asynchronous upload on its own stream into a pooled block)
    # while step k runs; `serial` builds it inline, for contrast.
    fresh = None
    if not args.no_fresh_supervision:
        try:
            sups_np = [sup] + [synth.make_supervision(5000 + 1000 * rank + k, den, B, T, num_paths=4, weight=1.0)
                               for k in range(3)]
            side = torch.cuda.Stream(dev)
            pool = ThreadPoolExecutor(max_workers=1)

            def build(k):
                torch.cuda.set_device(local)
                return R.new_supervision(sups_np[k % len(sups_np)]).prefetch(local, side)

            n_fresh = max(4, min(args.steps, 20))
            for mode in ("overlapped", "serial"):
                for timed in (False, True):
                    k, fut = 0, (pool.submit(build, 0) if mode == "overlapped" else None)
                    sync_all()
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ev0.record(stream)
                    for _ in range(n_fresh if timed else 2):
                        if mode == "overlapped":
                            h, fut = fut.result(), pool.submit(build, k + 1)
                        else:
                            h = R.new_supervision(sups_np[k % len(sups_np)])
                        step(s=h)
                        del h
                        k += 1
                    ev1.record(stream)
                    sync_all()
                    if mode == "overlapped":
                        fut.result()
                t_ms = R.max_over_ranks(ev0.elapsed_time(ev1)) / n_fresh
                fresh = dict(fresh or {}, **{mode + "_ms_per_step": t_ms})
            pool.shutdown()
            lib.torchain_supervision_pool_trim()
        except Exception as e:  # the headline line must survive a failure of an auxiliary measurement
            fresh = {"error": repr(e)[:200]}

    # ---- e2e: the reference-facing C-ABI call with HOST (pinned) buffers, copies inside the timed region
    y_pin = torch.empty(T * B, P, pin_memory=True)
    y_pin.copy_(y)
    d_pin = torch.empty(T * B, P, pin_memory=True)
    r_e2e = np.zeros(3, dtype=np.float32)

    def e2e_step():
        ok = lib.torchain_ComputeChainObjfAndDeriv_host(
            g.ptr, s.ptr, y_pin.data_ptr(), T * B, P, P, r_e2e.ctypes.data_as(ctypes.c_void_p),
            d_pin.data_ptr(), P, None, 0, c["l2"], c["leaky"], 0.0, local)
        if ok != 1:
            raise RuntimeError(_lib.last_error())

    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / n_e2e
    clocks = sampler.stop() if rank == 0 else None   # sampled over the device-timed AND the e2e-timed loops
    e2e_s = R.max_over_ranks(e2e_s)
    e2e_value = world * B * T / e2e_s

    # ---- parity: this rank's shard against the fp64 oracle on the same inputs (outside the timed regions)
    parity = None
    parity_error = None
    if not args.no_parity:
        try:
            ok = lib.torchain_ComputeChainObjfAndDeriv_ex(
                g.ptr, s.ptr, y.data_ptr(), T * B, P, y.stride(0), res_host.ctypes.data_as(ctypes.c_void_p), None,
                deriv.data_ptr(), deriv.stride(0), None, 0, c["l2"], c["leaky"], 0.0, -1.0, local,
                ctypes.c_void_p(stream.cuda_stream))          # un-reduced results of this rank's own shard
            if ok != 1:
                raise RuntimeError(_lib.last_error())
            torch.cuda.synchronize(dev)
            parity = parity_check(c, sup, y.cpu().numpy(), res_host.copy(), deriv.cpu().numpy(),
                                  max(1, host_cores() // world))
            d_host = d_pin.numpy()             # the e2e entry's derivative (not negated) must be the same matrix
            parity["e2e_deriv_vs_device_rel"] = float(
                np.abs(d_host + deriv.cpu().numpy()).max() / max(np.abs(d_host).max(), 1e-30))
            parity["e2e_objf_rel_err"] = abs(float(r_e2e[0]) - parity["oracle_objf"]) / abs(parity["oracle_objf"])
            parity["ok"] = bool(parity["ok"] and parity["e2e_deriv_vs_device_rel"] < 1e-4 and parity["e2e_objf_rel_err"] < 1e-5)
            if world > 1:                      # worst rank; and the all-reduced totals against the sum of the oracles
                tot = torch.tensor([parity["oracle_objf"]], device=dev, dtype=torch.float64)
                dist.all_reduce(tot)
                parity["reduced_objf_rel_err"] = abs(float(step_results[0]) - float(tot.item())) / abs(float(tot.item()))
                parity["ok"] = bool(parity["ok"] and parity["reduced_objf_rel_err"] < 1e-5)
                keys = ["objf_rel_err", "l2_rel_err", "weight_rel_err", "deriv_rel_err", "e2e_deriv_vs_device_rel",
                        "e2e_objf_rel_err", "oracle_s"]
                tmax = torch.tensor([parity[k] for k in keys] + [0.0 if parity["ok"] else 1.0], device=dev,
                                    dtype=torch.float64)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                for k, v in zip(keys, tmax.tolist()):
                    parity[k] = v
                parity["ok"] = bool(tmax[-1].item() == 0.0)
                parity["ranks_checked"] = world
            parity["tolerances"] = "objf, l2_term 1e-5 relative; derivative 1e-4 of max |oracle derivative|"
            parity["oracle"] = "oracle/chain_oracle.cpp fp64 mode, every rank's own shard, max over ranks"
        except Exception as e:  # reported, never fatal to the line (a missing parity block is visible as such)
            parity, parity_error = None, repr(e)[:200]

    # ---- the other BASELINE.json configurations on this launch geometry (flat keys on the same line):
    # cfg 3 = 512 sequences in total, strong-split over the ranks; cfg 4 = large tree, 256 sequences per GPU
    extras = {}
    if not args.no_extra_configs and args.config == "chime5" and not args.batch and not args.frames:
        try:
            del R
            torch.cuda.empty_cache()
            if 512 % world == 0:
                e3 = extra_config(torch, lib, "chime5", 512 // world, 50, dev, local, world, rank)
                extras.update(cfg3_strong_total_seqs=512, cfg3_strong_seqs_per_gpu=512 // world,
                              cfg3_strong_ms_per_step=e3["ms_per_step"], cfg3_strong_frames_per_sec=e3["frames_per_sec"],
                              cfg3_strong_gather_gbs_per_gpu=e3["gather_gbs"])
            e4 = extra_config(torch, lib, "large_tree", 0, 0, dev, local, world, rank)
            extras.update(cfg4_seqs_per_gpu=256, cfg4_ms_per_step=e4["ms_per_step"],
                          cfg4_frames_per_sec=e4["frames_per_sec"], cfg4_gather_gbs_per_gpu=e4["gather_gbs"])
            if world == 8:   # cfg 5's small end: 64 sequences over 8 GPUs
                e5 = extra_config(torch, lib, "chime5", 8, 50, dev, local, world, rank)
                extras.update(cfg5_64seqs_ms_per_step=e5["ms_per_step"], cfg5_64seqs_frames_per_sec=e5["frames_per_sec"])
        except Exception as e:
            extras["extra_configs_error"] = repr(e)[:200]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel pair (persistent denominator forward + backward launch)
    # The recurrence is an L2-gather: its binding level is the L2, not HBM, so the block's
    # achieved / peak / frac are the L2-gather figures and the HBM figures ride along as flat hbm_* keys.
    # Two measured denominators, both live in this run: a random 512 B-row gather over a 32 MB table
    # (the generic L2 figure) and the same gather over tables of this kernel's real sizes (5 MB state
    # rows + 1.5 MB exp rows, 32 warps per SM: the best point of the pattern sweep).  frac uses the LARGER.
    hbm_peak, peak_src = peaks()
    hbm_alg, gather_alg = alg_bytes(c)
    den_launch_ms = den_ms.value / max(1, den_n.value)
    hbm_gbs = hbm_alg * B * T / (den_launch_ms * 1e-3) / 1e9
    gather_gbs = gather_alg * B * T / (den_launch_ms * 1e-3) / 1e9
    sys.path.insert(0, os.path.join(ROOT, "tools", "microbench"))
    import mb as _mb
    mb = _mb.load()
    l2_peak_32mb = mb.tb_microbench_l2_gather(local, 32 << 20, 512, 400)
    l2_peak_tables = mb.tb_microbench_pattern(local, 32, 8, 3000)
    l2_peak = max(l2_peak_32mb, l2_peak_tables)
    traffic = measured_traffic() if args.config == "chime5" and not args.batch and not args.frames else None
    roofline = {
        "bound": "l2_gather", "kernel": "den_forward_kernel + den_backward_kernel (one launch pair per step)",
        "achieved": gather_gbs, "peak": l2_peak, "unit": "GB/s", "frac": gather_gbs / l2_peak if l2_peak > 0 else None,
        "peak_source": "measured live in this run (tools/microbench): max of the two gather peaks below",
        "gather_peak_32mb_table_gbs": l2_peak_32mb, "gather_peak_real_tables_gbs": l2_peak_tables,
        "frac_of_32mb_table_peak": gather_gbs / l2_peak_32mb if l2_peak_32mb > 0 else None,
        "frac_of_real_tables_peak": gather_gbs / l2_peak_tables if l2_peak_tables > 0 else None,
        "alg_bytes_per_launch": gather_alg * B * T, "launch_ms": den_launch_ms, "launches_timed": int(den_n.value),
        "hbm_gbs": hbm_gbs, "hbm_peak_gbs": hbm_peak, "hbm_frac": hbm_gbs / hbm_peak,
        "hbm_peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)", "hbm_alg_bytes_per_launch": hbm_alg * B * T,
        "traffic": (traffic or {}).get("dram_bytes_per_step"), "traffic_source": (traffic or {}).get("source"),
        "l2_traffic": (traffic or {}).get("l2_bytes_per_step"),
        "note": "achieved = 16*A algorithmic gather bytes per sequence-frame x B*T / launch-pair time (CUDA events "
                "inside the library); traffic = ncu dram bytes per launch pair, l2_traffic = ncu lts__t_sectors x 32",
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        fps, cores, sample = cpu_baseline(c)
        fps1, _, sample1 = cpu_baseline(c, seconds_budget=8.0, threads=1)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "value_1thread": fps1, "sample_1thread": sample1}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "chime5-shaped chain loss: P=%d N=%d A=%d, B=%d seqs x T=%d frames per GPU, leaky=%g l2=%g" % (
            P, c["N"], den.num_arcs, B, T, c["leaky"], c["l2"]),
            "name": args.config, "parallelism": "seq-shard x%d" % world,
            "l2_policy": "working set per step (y+deriv+alpha store = %.2f GB) exceeds the 126 MB L2; no flush" % (
                (2 * T * B * P * 4 + (T + 1) * c["N"] * B * 4) / 1e9)},
        "objf": float(step_results[0]), "l2_term": float(step_results[1]), "weight": float(step_results[2]),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": T * B * P * 4,
                "d2h_bytes_per_step": T * B * P * 4 + 12, "ms_per_step": e2e_s * 1e3,
                "api": "torchain_ComputeChainObjfAndDeriv_host (pinned host nnet_output in, derivative + results out)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if fresh and "error" in fresh:
        line["fresh_supervision_error"] = fresh["error"]
    elif fresh:
        line["fresh_supervision_ms_per_step"] = fresh["overlapped_ms_per_step"]
        line["fresh_supervision_serial_ms_per_step"] = fresh["serial_ms_per_step"]
        line["fresh_supervision_vs_step"] = fresh["overlapped_ms_per_step"] / ms_per_step
    line.update(extras)
    if cpu:
        line["cpu_baseline"] = cpu
    if parity_error:
        line["parity_error"] = parity_error
    if parity:
        line["parity"] = parity
        # flat copies at top level and inside `config` (which survives the driver's parser whole)
        line.update(parity_ok=parity["ok"], parity_objf_rel_err=parity["objf_rel_err"],
                    parity_deriv_rel_err=parity["deriv_rel_err"])
        line["config"].update(parity_ok=parity["ok"], parity_objf_rel_err=parity["objf_rel_err"],
                              parity_deriv_rel_err=parity["deriv_rel_err"])
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_OUT_FD = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries print there too (NCCL writes
    "NCCL version ..." with printf during init): keep the real stdout for the result line and point
    file descriptor 1 at stderr for everything else."""
    global _OUT_FD
    sys.stdout.flush()
    _OUT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _OUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_OUT_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="chime5")
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU (default: the config's)")
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the fp64-oracle check of the step's results")
    ap.add_argument("--no-fresh-supervision", action="store_true", help="skip the new-Supervision-per-step loops")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip cfg 3 (strong split) / cfg 4 extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
