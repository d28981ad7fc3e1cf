"""Host logic of libtorchain_b200 end to end on the CPU.

tests/mock_cuda links the library's OWN object files against a host-memory stand-in for the CUDA runtime
(device memory = host memory, copies = memcpy, kernel launches only recorded).  Nothing is computed -- the
arithmetic is the GPU suite's business -- but everything the host does around the kernels runs for real:
handle uploads, workspace management, the launch sequence and geometry, the chunk loops of the host-buffer
entry.  The mock is test infrastructure; the product library never links it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from torchain_b200 import _lib, synth

HERE = os.path.dirname(os.path.abspath(__file__))
MOCK_DIR = os.path.join(HERE, "mock_cuda")
MOCK_LIB = os.path.join(MOCK_DIR, "libtorchain_b200_mock.so")


class NumParams(ctypes.Structure):  # torchain_b200/csrc/chain_kernels.h, tb::NumParams
    _fields_ = [("num_chains", ctypes.c_int)] + \
               [(n, ctypes.c_void_p) for n in ("chain_level_off", "level_state_off", "chain_arc_off", "in_off", "in_arcs",
                                               "out_off", "out_arcs", "final_cost", "arc_row", "arc_pdf", "arc_cost")] + \
               [("B", ctypes.c_int), ("y", ctypes.c_void_p), ("yl", ctypes.c_int64 * 3)] + \
               [(n, ctypes.c_void_p) for n in ("arc_val", "log_alpha", "log_beta", "arc_post", "chain_logprob")]


class DenDir(ctypes.Structure):  # tb::DenDir
    _fields_ = [(n, ctypes.c_void_p) for n in ("arcs", "tasks", "cta_arc_off", "cta_task_off", "warp_task_off",
                                               "warp_arc_off")] + [("max_arcs_cta", ctypes.c_int), ("max_tasks_cta", ctypes.c_int)]


class DenParams(ctypes.Structure):  # tb::DenParams
    _fields_ = [("dir", DenDir * 2)] + \
               [(n, ctypes.c_int) for n in ("N", "P", "B", "BP", "Btot", "seq0", "T")] + \
               [("y", ctypes.c_void_p), ("yl", ctypes.c_int64 * 3), ("deriv", ctypes.c_void_p), ("dl", ctypes.c_int64 * 3)] + \
               [(n, ctypes.c_void_p) for n in ("alpha", "beta", "etile", "gtile", "Ssum", "Bsum", "scalars", "barrier")] + \
               [(n, ctypes.c_float) for n in ("leaky", "sum_pi", "scale_gamma", "scale_y")] + \
               [(n, ctypes.c_int) for n in ("arcs_in_smem", "subs", "has_split")] + \
               [(n, ctypes.c_void_p) for n in ("num_frame_off", "num_frame_arcs", "num_arc_pdf", "num_arc_seq",
                                               "num_arc_post", "y_ready")] + \
               [("y_chunk_frames", ctypes.c_int), ("deriv_progress", ctypes.c_void_p), ("progress_mask", ctypes.c_uint64 * 8)] + \
               [("rdir", DenDir), ("R", ctypes.c_int), ("arun", ctypes.c_void_p), ("wbuf", ctypes.c_void_p)] + \
               [(n, ctypes.c_void_p) for n in ("warp_busy", "cta_trace", "timeline")]


class DenBatch(ctypes.Structure):  # tb::DenBatch
    _fields_ = [("c", DenParams * 2), ("nctx", ctypes.c_int)]


@pytest.fixture(scope="module")
def mock():
    objs = os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", "chain_api.o")
    if not os.path.exists(objs):
        pytest.skip("library objects not built (run __graft_entry__.build())")
    path = os.environ.get("TORCHAIN_MOCK_LIB")   # e.g. a sanitizer build of the host code
    if not path:
        subprocess.run(["make", "-C", MOCK_DIR, "-s"], check=True)
        path = MOCK_LIB
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, argtypes
    lib.mock_launch_name.restype = ctypes.c_char_p
    lib.mock_launch_param0.restype = ctypes.c_void_p
    lib.mock_stat.restype = ctypes.c_long
    lib.mock_set_param_capture.argtypes = [ctypes.c_size_t]
    lib.mock_set_param_capture_for.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    lib.mock_set_param_capture_for(b"num_fb_kernel", ctypes.sizeof(NumParams))
    lib.mock_set_param_capture_for(b"den_forward_kernel", ctypes.sizeof(DenBatch))
    lib.mock_set_param_capture_for(b"den_backward_kernel", ctypes.sizeof(DenBatch))
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _handles(lib, den, sup):
    keep = [np.ascontiguousarray(x) for x in (den.src, den.dst, den.ilabel, den.cost, den.final_cost, sup.arc_begin,
                                              sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)]
    g = lib.torchain_denominator_graph_from_arrays(den.num_states, den.num_arcs, _p(keep[0]), _p(keep[1]), _p(keep[2]),
                                                   _p(keep[3]), den.start, _p(keep[4]), den.num_pdfs)
    s = lib.torchain_supervision_from_arrays(sup.weight, sup.num_sequences, sup.frames_per_sequence, sup.label_dim,
                                             sup.num_states, _p(keep[5]), _p(keep[6]), _p(keep[7]), _p(keep[8]),
                                             _p(keep[9]))
    assert g and s, lib.torchain_last_error()
    return g, s


def _compute(lib, g, s, y, stream=None):
    d = np.full_like(y, np.nan)
    res = np.full(3, np.nan, dtype=np.float32)
    ok = lib.my_lib_ComputeChainObjfAndDeriv(g, s, _p(y), y.shape[0], y.shape[1], y.shape[1], _p(res), _p(d), y.shape[1],
                                             None, 0, 5e-5, 0.1, 0.0, 0, stream)
    assert ok == 1, lib.torchain_last_error()
    return res, d


def _launches(lib):
    out = []
    for i in range(lib.mock_launch_count()):
        grid, block, coop, smem = ctypes.c_uint(), ctypes.c_uint(), ctypes.c_int(), ctypes.c_ulonglong()
        lib.mock_launch_info(i, ctypes.byref(grid), ctypes.byref(block), ctypes.byref(coop), None, ctypes.byref(smem))
        out.append((lib.mock_launch_name(i).decode(), grid.value, block.value, coop.value, smem.value))
    return out


def _kernel(name):
    for k in ("num_fb_kernel", "den_tied_forward_kernel", "den_tied_backward_kernel", "den_forward_kernel",
              "den_backward_kernel", "final_scalars_kernel", "final_deriv_kernel",
              "zero_kernel"):
        if k in name:
            return k
    return name


@pytest.mark.parametrize("B,subs,passes", [(16, 4, [1]), (64, 2, [1]), (128, 1, [1]), (200, 1, [2]), (300, 1, [2, 1])])
def test_launch_sequence_and_geometry(mock, B, subs, passes):
    """One call = numerator (side stream) + one forward / backward launch pair per pass of <= 256 sequences +
    the two final kernels; the denominator launches are cooperative, one CTA of 512 threads per SM, and their
    template instantiation follows the batch geometry (lane groups per warp, one or two contexts)."""
    den = synth.make_den_fst(1, 300, 40, mean_deg=4.0)
    T = 5
    sup = synth.make_supervision(2, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(3, B, T, den.num_pdfs).astype(np.float32)
    g, s = _handles(mock, den, sup)
    try:
        mock.mock_reset_launches()
        n0 = mock.torchain_kernel_launch_count()
        _compute(mock, g, s, y)
        ls = _launches(mock)
        assert mock.torchain_kernel_launch_count() - n0 == len(ls) == 3 + 2 * len(passes)
        assert [_kernel(l[0]) for l in ls] == (["num_fb_kernel"] + ["den_forward_kernel", "den_backward_kernel"] * len(passes) +
                                               ["final_scalars_kernel", "final_deriv_kernel"])
        den_ls = [l for l in ls if "den_" in l[0]]
        for (name, grid, block, coop, smem), nctx in zip(den_ls, [c for c in passes for _ in range(2)]):
            assert (grid, block, coop) == (148, 512, 1)
            sub_of_pass = subs if len(passes) == 1 else (1 if nctx == 2 else 2)   # second pass of B=300: 44 -> BP 64
            assert "kernelILi%dELb1ELb0ELi%dEE" % (sub_of_pass, nctx) in name     # <SUBS, STAGED, SPLIT, NCTX>
            assert 0 < smem <= 227 * 1024
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_supervision_device_image_is_one_pooled_block_with_the_right_contents(mock):
    """The device image of a supervision is ONE block (device + pinned staging) out of a process-wide pool:
    at most one device allocation when it is first used, no cudaFree when the handle dies (a training loop
    builds a supervision per minibatch, and cudaFree synchronises the device), and the next supervision
    reuses the block."""
    den = synth.make_den_fst(4, 80, 20, mean_deg=3.0)
    B, T = 1, 9      # one sequence: the device arcs are the FST's arcs one to one
    sup = synth.make_supervision(5, den, B, T, num_paths=3, weight=1.0)
    y = synth.make_nnet_output(6, B, T, den.num_pdfs).astype(np.float32)
    g, s = _handles(mock, den, sup)
    try:
        _compute(mock, g, s, y)                      # builds the graph's device state and the workspace
        s2 = _handles(mock, den, sup)[1]
        m0, f0 = mock.mock_stat(0), mock.mock_stat(1)
        mock.mock_reset_launches()
        _compute(mock, g, s2, y)                     # only the new supervision has to be uploaded
        assert mock.mock_stat(0) - m0 <= 1           # at most ONE device allocation (none if the pool had a block) ...
        names = [mock.mock_launch_name(i).decode() for i in range(mock.mock_launch_count())]
        i_num = [i for i, n in enumerate(names) if "num_fb_kernel" in n][0]
        npar = NumParams.from_address(mock.mock_launch_param0(i_num))
        A, S = sup.num_arcs, sup.num_states
        arr = lambda ptr, ct, n: np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,)).copy()
        assert npar.num_chains == 1 and npar.B == B
        assert list(arr(npar.chain_arc_off, ctypes.c_int, 2)) == [0, A]
        in_off, out_off = arr(npar.in_off, ctypes.c_int, S + 1), arr(npar.out_off, ctypes.c_int, S + 1)
        assert in_off[0] == 0 and in_off[-1] == A and (np.diff(in_off) >= 0).all()
        assert np.array_equal(out_off, sup.arc_begin)
        # time of an arc's source state = frame (one sequence): row = frame * B + 0
        time = np.zeros(S, dtype=np.int64)
        for st in range(S):
            for a in range(sup.arc_begin[st], sup.arc_begin[st + 1]):
                time[sup.arc_dst[a]] = time[st] + 1
        src = np.repeat(np.arange(S), np.diff(sup.arc_begin))
        assert np.array_equal(arr(npar.arc_row, ctypes.c_int, A), time[src] * B)
        assert np.array_equal(arr(npar.arc_pdf, ctypes.c_int, A), sup.arc_label - 1)
        assert np.array_equal(arr(npar.arc_cost, ctypes.c_float, A), sup.arc_cost)
        assert np.array_equal(arr(npar.final_cost, ctypes.c_float, S), sup.final_cost)
        out_arcs = arr(npar.out_arcs, ctypes.c_int, 2 * A).reshape(A, 2)
        assert np.array_equal(out_arcs[:, 0], sup.arc_dst) and np.array_equal(out_arcs[:, 1], np.arange(A))
        in_arcs = arr(npar.in_arcs, ctypes.c_int, 2 * A).reshape(A, 2)
        assert np.array_equal(np.sort(in_arcs[:, 1]), np.arange(A)) and np.array_equal(in_arcs[:, 0], src[in_arcs[:, 1]])
        assert np.array_equal(arr(npar.y, ctypes.c_float, y.size).reshape(y.shape), y)
        mock.my_lib_supervision_free(s2)
        assert mock.mock_stat(1) == f0               # ... NO free: the block went back to the pool
        m1 = mock.mock_stat(0)
        s3 = _handles(mock, den, sup)[1]
        _compute(mock, g, s3, y)
        assert mock.mock_stat(0) == m1               # ... and the next minibatch's supervision reuses it
        mock.my_lib_supervision_free(s3)
        assert mock.mock_stat(1) == f0
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_workspaces_are_capped_per_device(mock):
    """A caller hopping between streams (torch pools 32 per device) keeps at most 4 workspaces alive."""
    den = synth.make_den_fst(7, 120, 16, mean_deg=3.0)
    B, T = 8, 6
    sup = synth.make_supervision(8, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(9, B, T, den.num_pdfs).astype(np.float32)
    g, s = _handles(mock, den, sup)
    try:
        streams0 = mock.mock_stat(5)
        live = []
        for k in range(1, 13):
            _compute(mock, g, s, y, ctypes.c_void_p(0x1000 * k))   # stream handles are opaque tokens to the mock
            live.append((mock.mock_stat(5) - streams0, mock.mock_stat(2)))
        assert [n for n, _ in live] == [1, 2, 3, 4] + [4] * 8       # one side stream per workspace
        assert len({b for _, b in live[3:]}) == 1                   # device memory stops growing
        _compute(mock, g, s, y, ctypes.c_void_p(0x1000 * 12))        # the most recent one is still there
        assert mock.mock_stat(5) - streams0 == 4
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)
    assert mock.mock_stat(5) == streams0


@pytest.mark.parametrize("progress", [False, True], ids=["late-output", "pipelined-output"])
@pytest.mark.parametrize("T,pad", [(1, 0), (3, 0), (17, 4), (50, 0), (77, 3), (130, 1)])
def test_host_buffer_entry_moves_every_row_once(mock, progress, T, pad):
    """Chunk loops of the host-buffer entry for every schedule (chunks alternate between two copy streams), with and
    without the backward launch's progress counter advancing while the host polls it: all of nnet_output reaches the
    device buffer, every row of the derivative (zeros here: no kernel ran) reaches the host buffer, the padding of
    pitched buffers is left alone."""
    mock.mock_set_auto_progress(T if progress else 0)
    den = synth.make_den_fst(10, 60, 12, mean_deg=3.0)
    B = 5
    sup = synth.make_supervision(11, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(12, B, T, den.num_pdfs).astype(np.float32)
    rows, cols = y.shape
    yp = np.full((rows, cols + pad), 7.0, dtype=np.float32)
    yp[:, :cols] = y
    dp = np.full((rows, cols + pad), np.nan, dtype=np.float32)
    res = np.full(3, np.nan, dtype=np.float32)
    g, s = _handles(mock, den, sup)
    try:
        mock.mock_reset_launches()
        ok = mock.torchain_ComputeChainObjfAndDeriv_host(g, s, _p(yp), rows, cols, cols + pad, _p(res), _p(dp),
                                                         cols + pad, None, 0, 5e-5, 0.1, 0.0, 0)
        assert ok == 1, mock.torchain_last_error()
        assert np.all(dp[:, :cols] == 0.0) and np.isnan(dp[:, cols:]).all()
        assert np.all(res == 0.0)
        names = [mock.mock_launch_name(i).decode() for i in range(mock.mock_launch_count())]
        npar = NumParams.from_address(mock.mock_launch_param0([i for i, n in enumerate(names) if "num_fb_kernel" in n][0]))
        y_dev = np.ctypeslib.as_array(ctypes.cast(npar.y, ctypes.POINTER(ctypes.c_float)), shape=(rows, cols))
        assert np.array_equal(y_dev, y)
    finally:
        mock.mock_set_auto_progress(0)
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_host_buffer_entry_redelivers_early_chunks_when_the_guard_tripped(mock):
    """With the mock standing in for the backward kernel's progress counter (one frame per stream query, up to
    the T-2 the kernel publishes) the chunks of frames >= 2 leave early; the mock's final pass reports
    ok = 0 (zeroed results), i.e. a tripped numeric guard, so exactly those chunks must be copied a second time."""
    den = synth.make_den_fst(10, 60, 12, mean_deg=3.0)
    B, T = 5, 50
    sup = synth.make_supervision(11, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(12, B, T, den.num_pdfs).astype(np.float32)
    rows, cols = y.shape
    res = np.zeros(3, dtype=np.float32)
    g, s = _handles(mock, den, sup)
    n = mock.torchain_host_chunk_bounds(T, None, 0)
    bounds = (ctypes.c_int * n)()
    mock.torchain_host_chunk_bounds(T, bounds, n)
    early = sum(1 for k in range(n - 1) if T - bounds[k] <= T - 2)   # chunks whose frames are all >= 2

    def call():
        d = np.full_like(y, np.nan)
        c0 = mock.mock_stat(3)
        ok = mock.torchain_ComputeChainObjfAndDeriv_host(g, s, _p(y), rows, cols, cols, _p(res), _p(d), cols, None, 0,
                                                         5e-5, 0.1, 0.0, 0)
        assert ok == 1, mock.torchain_last_error()
        assert np.all(d == 0.0)
        return mock.mock_stat(3) - c0

    try:
        call()                       # one-off uploads
        plain = call()               # the stream "finishes" at once: nothing leaves early
        mock.mock_set_auto_progress(T - 2)
        try:
            piped = call()
        finally:
            mock.mock_set_auto_progress(0)
        assert early == 15 and piped - plain == early
        # ... and T-1 after the kernel's last grid barrier: the single-frame chunk of frame 1 leaves early too
        mock.mock_set_auto_progress(T - 1)
        try:
            assert call() - plain == early + 1 == n - 2
        finally:
            mock.mock_set_auto_progress(0)
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_no_device_memory_or_stream_outlives_the_handles(mock):
    den = synth.make_den_fst(13, 150, 24, mean_deg=4.0)
    B, T = 40, 7
    sup = synth.make_supervision(14, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(15, B, T, den.num_pdfs).astype(np.float32)
    mock.torchain_supervision_pool_trim()
    live0, streams0, balance0 = mock.mock_stat(2), mock.mock_stat(5), mock.mock_stat(0) - mock.mock_stat(1)
    g, s = _handles(mock, den, sup)
    for k in range(3):
        _compute(mock, g, s, y, ctypes.c_void_p(0x2000 * (1 + k % 2)))
    assert mock.mock_stat(2) > live0
    mock.my_lib_supervision_free(s)
    mock.my_lib_denominator_graph_free(g)
    mock.torchain_supervision_pool_trim()
    assert mock.mock_stat(2) == live0 and mock.mock_stat(5) == streams0
    assert mock.mock_stat(0) - mock.mock_stat(1) == balance0


def test_host_threads_share_the_handles(mock):
    """The reference drives several GPUs / streams from Python threads (parallel_train.py:55-75); ctypes drops
    the GIL, so these calls really run concurrently: shared den graph, own stream, a new supervision per call."""
    import threading
    den = synth.make_den_fst(16, 200, 30, mean_deg=4.0)
    B, T = 12, 6
    sup = synth.make_supervision(17, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(18, B, T, den.num_pdfs).astype(np.float32)
    mock.torchain_supervision_pool_trim()
    live0, streams0 = mock.mock_stat(2), mock.mock_stat(5)
    g, s0 = _handles(mock, den, sup)
    n0 = mock.torchain_kernel_launch_count()
    errors = []

    def worker(tid):
        try:
            for it in range(40):
                s = _handles(mock, den, sup)[1] if it % 2 else None
                _compute(mock, g, s or s0, y, ctypes.c_void_p(0x3000 * (1 + tid)))
                if s:
                    mock.my_lib_supervision_free(s)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert mock.torchain_kernel_launch_count() - n0 == 6 * 40 * 5
    assert mock.mock_stat(5) - streams0 <= 6      # 6 streams in use at once: never more workspaces than callers
    mock.my_lib_supervision_free(s0)
    mock.my_lib_denominator_graph_free(g)
    mock.torchain_supervision_pool_trim()
    assert mock.mock_stat(2) == live0 and mock.mock_stat(5) == streams0


def test_two_devices_in_one_process(mock):
    """parallel_train.py drives several GPUs from one process: the den-graph handle keeps one partition upload
    and one workspace per device, and a call leaves the caller's current device as it found it."""
    den = synth.make_den_fst(19, 100, 16, mean_deg=3.0)
    B, T = 8, 5
    sup = synth.make_supervision(20, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(21, B, T, den.num_pdfs).astype(np.float32)
    mock.mock_set_device_count(2)
    g, s = _handles(mock, den, sup)
    try:
        d = np.empty_like(y)
        res = np.zeros(3, dtype=np.float32)
        grown = []
        for dev in (0, 1, 1, 0):
            m0 = mock.mock_stat(0)
            ok = mock.my_lib_ComputeChainObjfAndDeriv(g, s, _p(y), y.shape[0], y.shape[1], y.shape[1], _p(res), _p(d),
                                                      y.shape[1], None, 0, 0.0, 0.1, 0.0, dev, None)
            assert ok == 1, mock.torchain_last_error()
            grown.append(mock.mock_stat(0) - m0)
            cur = ctypes.c_int(-1)
            mock.cudaGetDevice(ctypes.byref(cur))
            assert cur.value == 0
        assert grown[0] > 0 and grown[1] > 0 and grown[2] == 0 and grown[3] == 0   # uploads once per device
        assert mock.my_lib_ComputeChainObjfAndDeriv(g, s, _p(y), y.shape[0], y.shape[1], y.shape[1], _p(res), _p(d),
                                                    y.shape[1], None, 0, 0.0, 0.1, 0.0, 2, None) == 0
        assert b"device" in mock.torchain_last_error()
    finally:
        mock.mock_set_device_count(1)
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_partition_autotuner_redeals_and_keeps_old_generations(mock):
    """Synchronising full-width calls feed the partition autotuner (per-warp busy counters -> speeds -> LPT).
    With the mock handing out pre-filled "device" memory the counters read as large equal numbers, so every one
    of the three calibration rounds builds and uploads a new partition generation; after that the partition is
    frozen (no further allocations), and every launch still gets a valid geometry."""
    den = synth.make_den_fst(22, 3000, 64, mean_deg=6.0)
    B, T = 128, 3
    sup = synth.make_supervision(23, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(24, B, T, den.num_pdfs).astype(np.float32)
    mock.torchain_supervision_pool_trim()
    live0 = mock.mock_stat(2)
    mock.mock_set_fill(1)
    g, s = _handles(mock, den, sup)
    try:
        grown = []
        for _ in range(6):
            m0 = mock.mock_stat(0)
            mock.mock_reset_launches()
            _compute(mock, g, s, y)
            grown.append(mock.mock_stat(0) - m0)
            for name, grid, block, coop, smem in _launches(mock):
                if "den_" in name:
                    assert (grid, block, coop) == (148, 512, 1) and 0 < smem <= 227 * 1024
        assert grown[0] > grown[1] > 0 and grown[1] == grown[2] == 12   # one generation = 2 directions x 6 arrays
        assert grown[3:] == [0, 0, 0]
    finally:
        mock.mock_set_fill(0)
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)
    mock.torchain_supervision_pool_trim()
    assert mock.mock_stat(2) == live0


def _den_batches(lib):
    out = []
    for i in range(lib.mock_launch_count()):
        name = lib.mock_launch_name(i).decode()
        if "den_forward_kernel" in name or "den_backward_kernel" in name:
            out.append(("fwd" if "forward" in name else "bwd", DenBatch.from_buffer_copy(
                ctypes.string_at(lib.mock_launch_param0(i), ctypes.sizeof(DenBatch)))))
    return out


def test_denominator_launch_parameters(mock):
    """What the host hands the persistent kernels: geometry, sequence ranges of the contexts and passes,
    per-context buffers, the scales that fold supervision.weight / l2 / the sign into the derivative."""
    assert ctypes.sizeof(DenParams) == 536 and ctypes.sizeof(DenBatch) == 1080   # mirrors of chain_kernels.h
    den = synth.make_den_fst(25, 500, 48, mean_deg=5.0)
    B, T, l2, leaky, w = 300, 4, 0.01, 0.2, 0.5
    sup = synth.make_supervision(26, den, B, T, num_paths=2, weight=w)
    y = synth.make_nnet_output(27, B, T, den.num_pdfs).astype(np.float32)
    g, s = _handles(mock, den, sup)
    try:
        d = np.empty_like(y)
        res = np.zeros(3, dtype=np.float32)
        mock.mock_reset_launches()
        ok = mock.torchain_ComputeChainObjfAndDeriv_ex(g, s, _p(y), y.shape[0], y.shape[1], y.shape[1], _p(res), None,
                                                       _p(d), y.shape[1], None, 0, l2, leaky, 0.0, -1.0, 0, None)
        assert ok == 1, mock.torchain_last_error()
        batches = _den_batches(mock)
        assert [k for k, _ in batches] == ["fwd", "bwd", "fwd", "bwd"]
        (_, f1), (_, b1), (_, f2), (_, b2) = batches
        assert bytes(f1) == bytes(b1) and bytes(f2) == bytes(b2)          # a launch pair shares its parameters
        # pass 1: 256 sequences as two contexts of 128; pass 2: the remaining 44, one context padded to 64
        assert f1.nctx == 2 and [(c.seq0, c.B, c.BP, c.subs) for c in f1.c] == [(0, 128, 128, 1), (128, 128, 128, 1)]
        assert f2.nctx == 1 and (f2.c[0].seq0, f2.c[0].B, f2.c[0].BP, f2.c[0].subs) == (256, 44, 64, 2)
        for bt in (f1, f2):
            for c in bt.c[:bt.nctx]:
                assert (c.N, c.P, c.Btot, c.T) == (den.num_states, den.num_pdfs, B, T)
                assert c.y == y.ctypes.data and c.deriv == d.ctypes.data
                assert list(c.yl) == [B * y.shape[1], y.shape[1], 1] and list(c.dl) == list(c.yl)
                assert c.leaky == pytest.approx(leaky) and c.arcs_in_smem == 1 and c.has_split == 0
                assert c.scale_gamma == pytest.approx(-1.0 * -w) and c.scale_y == pytest.approx(-1.0 * -w * l2)
                assert c.y_ready is None and c.deriv_progress is None
                for d_ in c.dir:
                    assert d_.arcs and d_.tasks and d_.max_arcs_cta > 0 and d_.max_arcs_cta % 4 == 0
        a, b = f1.c
        for name in ("alpha", "beta", "etile", "gtile", "Ssum", "Bsum", "barrier"):
            assert getattr(a, name) != getattr(b, name)                   # own buffers per context
        assert a.scalars == b.scalars                                     # one set of sums per call
        # alpha store of a context: (T + 1) x N x BP floats apart
        assert b.alpha - a.alpha == 4 * (T + 1) * den.num_states * 128
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_host_buffer_entry_launch_parameters(mock):
    den = synth.make_den_fst(28, 80, 12, mean_deg=3.0)
    B, T = 6, 20
    sup = synth.make_supervision(29, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(30, B, T, den.num_pdfs).astype(np.float32)
    d = np.empty_like(y)
    res = np.zeros(3, dtype=np.float32)
    g, s = _handles(mock, den, sup)
    try:
        mock.mock_reset_launches()
        ok = mock.torchain_ComputeChainObjfAndDeriv_host(g, s, _p(y), y.shape[0], y.shape[1], y.shape[1], _p(res), _p(d),
                                                         y.shape[1], None, 0, 0.0, 0.1, 0.0, 0)
        assert ok == 1, mock.torchain_last_error()
        (_, f), (_, b) = _den_batches(mock)
        c = f.c[0]
        assert c.y != y.ctypes.data and c.deriv != d.ctypes.data          # device copies, not the host buffers
        assert c.y_ready and c.y_chunk_frames == 1 and c.deriv_progress   # one flag per frame, progress counter
        flags = np.ctypeslib.as_array(ctypes.cast(c.y_ready, ctypes.POINTER(ctypes.c_uint)), shape=(T,))
        assert (flags != 0).all()                                          # every frame was announced
        assert c.scale_gamma == pytest.approx(-1.0) and c.scale_y == 0.0   # deriv_scale = +1 on this entry
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_strided_xent_and_objective_only_entries(mock):
    den = synth.make_den_fst(31, 90, 14, mean_deg=3.0)
    B, T = 7, 9
    P = den.num_pdfs
    sup = synth.make_supervision(32, den, B, T, num_paths=2, weight=1.0)
    g, s = _handles(mock, den, sup)
    try:
        # a network output left in (B, P, T): element strides over (frame, sequence, pdf) = (1, P*T, T)
        x = np.zeros((B, P, T), dtype=np.float32)
        dx = np.zeros_like(x)
        xe = np.zeros_like(x)
        st = (ctypes.c_int64 * 3)(1, P * T, T)
        res = np.zeros(3, dtype=np.float32)
        mock.mock_reset_launches()
        ok = mock.torchain_ComputeChainObjfAndDeriv_strided(g, s, _p(x), T, B, P, st, _p(res), None, _p(dx), st, _p(xe), st,
                                                            0.0, 0.1, 0.1, -1.0, 0, None)
        assert ok == 1, mock.torchain_last_error()
        kernels = [_kernel(lib_name) for lib_name, *_ in _launches(mock)]
        assert kernels == ["num_fb_kernel", "den_forward_kernel", "den_backward_kernel", "zero_kernel",
                           "final_scalars_kernel", "final_deriv_kernel"]    # the xent output is zeroed, then scattered
        c = _den_batches(mock)[0][1].c[0]
        assert list(c.yl) == [1, P * T, T] and list(c.dl) == [1, P * T, T] and c.y == x.ctypes.data
        # frames / sequences that do not match the supervision, missing strides
        assert mock.torchain_ComputeChainObjfAndDeriv_strided(g, s, _p(x), T + 1, B, P, st, _p(res), None, _p(dx), st, None,
                                                              None, 0.0, 0.1, 0.0, -1.0, 0, None) == 0
        assert mock.torchain_ComputeChainObjfAndDeriv_strided(g, s, _p(x), T, B, P, st, _p(res), None, _p(dx), None, None,
                                                              None, 0.0, 0.1, 0.0, -1.0, 0, None) == 0
        # objective only: no derivative buffer from the caller -> a scratch matrix in the workspace
        y = np.zeros((T * B, P), dtype=np.float32)
        mock.mock_reset_launches()
        ok = mock.torchain_ComputeChainObjfAndDeriv_ex(g, s, _p(y), T * B, P, P, _p(res), None, None, 0, None, 0,
                                                       0.0, 0.1, 0.0, 1.0, 0, None)
        assert ok == 1, mock.torchain_last_error()
        c = _den_batches(mock)[0][1].c[0]
        assert c.deriv and c.deriv != y.ctypes.data and list(c.dl) == [B * P, P, 1]
        # leaky-HMM coefficient outside (0, 1): Kaldi asserts, this returns an error
        for bad in (0.0, 1.0, -0.1, float("nan")):
            assert mock.torchain_ComputeChainObjfAndDeriv_ex(g, s, _p(y), T * B, P, P, _p(res), None, None, 0, None, 0,
                                                             0.0, bad, 0.0, 1.0, 0, None) == 0
            assert b"leaky" in mock.torchain_last_error()
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_relayout_entries_launch_geometry(mock):
    """torchain_to2d / torchain_from2d: one CTA per (sequence, 32 pdfs) block; sizes the tiled kernel cannot
    take are refused so that the caller permutes some other way."""
    B, P, T = 5, 70, 12
    x = np.zeros((B, P, T), dtype=np.float32)
    rows = np.zeros((T * B, P), dtype=np.float32)
    mock.mock_reset_launches()
    assert mock.torchain_to2d(_p(x), B, P, T, _p(rows), P, 0, None) == 1
    assert mock.torchain_from2d(_p(rows), P, B, P, T, _p(x), -1.0, 0, None) == 1
    ls = _launches(mock)
    assert len(ls) == 2 and all("relayout" in l[0] or "2d" in l[0] for l in ls)
    assert mock.torchain_to2d(_p(x), B, P, 1600, _p(rows), P, 0, None) == 0
    assert mock.torchain_to2d(_p(x), B, P, T, _p(rows), P - 1, 0, None) == 0
    assert mock.torchain_from2d(None, P, B, P, T, _p(x), 1.0, 0, None) == 0


def test_host_buffer_entry_from_several_threads(mock):
    """Per-thread resources of the host-buffer entry (streams, device buffers, flags, progress counter):
    three host threads call it concurrently on shared handles."""
    import threading
    den = synth.make_den_fst(33, 70, 10, mean_deg=3.0)
    B, T = 4, 11
    sup = synth.make_supervision(34, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(35, B, T, den.num_pdfs).astype(np.float32)
    g, s = _handles(mock, den, sup)
    errors = []

    def worker():
        try:
            for _ in range(25):
                d = np.full_like(y, np.nan)
                res = np.full(3, np.nan, dtype=np.float32)
                ok = mock.torchain_ComputeChainObjfAndDeriv_host(g, s, _p(y), y.shape[0], y.shape[1], y.shape[1], _p(res),
                                                                 _p(d), y.shape[1], None, 0, 0.0, 0.1, 0.0, 0)
                assert ok == 1 and np.all(d == 0.0) and np.all(res == 0.0)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker) for _ in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    mock.my_lib_supervision_free(s)
    mock.my_lib_denominator_graph_free(g)
    assert not errors, errors


# ---------------------------------------------------------------------------------------------------------
# Run-based ("tied") kernels: the streams the host builds, executed in numpy.
# ---------------------------------------------------------------------------------------------------------
def _arr(ptr, ct, n):
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,)).copy()


def _dir_arrays(d, nbins, grid):
    warp_arc_off = _arr(d.warp_arc_off, ctypes.c_int, nbins + 1)
    out = dict(warp_arc_off=warp_arc_off, recs=_arr(d.arcs, ctypes.c_int32, 4 * int(warp_arc_off[-1])).reshape(-1, 4))
    if d.tasks:
        out["warp_task_off"] = _arr(d.warp_task_off, ctypes.c_int, nbins + 1)
        out["tasks"] = _arr(d.tasks, ctypes.c_int32, 4 * int(out["warp_task_off"][-1])).reshape(-1, 4)
        out["cta_arc_off"] = _arr(d.cta_arc_off, ctypes.c_int, grid + 1)
    return out


def _f32(i):
    return np.array([i], dtype=np.int32).view(np.float32)[0].astype(np.float64)


def test_tied_streams_executed_in_numpy_match_the_oracle(mock):
    """A graph with tied pdfs takes the run-based kernels (den_tied_*).  The mock records their launch parameters;
    this test walks the uploaded record streams exactly as the kernels do (forward: arc / run records with one row
    per record; backward phase A: runs; phase B: arcs pointing at their run's W row) in float64 for one sequence
    and compares log-likelihood and occupancies with the oracle's per-arc formulation."""
    mock.mock_set_param_capture_for(b"den_tied_forward_kernel", ctypes.sizeof(DenBatch))
    mock.mock_set_param_capture_for(b"den_tied_backward_kernel", ctypes.sizeof(DenBatch))
    P, N, T, leaky = 30, 150, 7, 0.1
    den = synth.make_den_fst_tied(5, N, P, mean_deg=8.0)
    runs_expected, _ = synth.in_pdf_runs(den)
    assert runs_expected * 3 <= den.num_arcs
    B = 65                                              # the tied path needs a pass of more than 64 sequences
    sup = synth.make_supervision(6, den, B, T, num_paths=2)
    y = synth.make_nnet_output(7, B, T, P, scale=1.5).astype(np.float32)
    g, s = _handles(mock, den, sup)
    try:
        mock.mock_reset_launches()
        _compute(mock, g, s, y)
        names = [mock.mock_launch_name(i).decode() for i in range(mock.mock_launch_count())]
        assert [_kernel(n) for n in names if "den_" in n] == ["den_tied_forward_kernel", "den_tied_backward_kernel"]
        i_f = [i for i, n in enumerate(names) if "den_tied_forward_kernel" in n][0]
        c = DenBatch.from_buffer_copy(ctypes.string_at(mock.mock_launch_param0(i_f), ctypes.sizeof(DenBatch))).c[0]
        grid, bins, rowbytes = 148, 14, 4 * 128
        assert (c.BP, c.subs, c.N, c.P, c.R) == (128, 1, N, P, runs_expected)
        fwd, bwd = _dir_arrays(c.dir[0], grid * bins, grid), _dir_arrays(c.dir[1], grid * bins, grid)
        run_off = _arr(c.rdir.warp_arc_off, ctypes.c_int, grid * bins + 1)
        runs = _arr(c.rdir.arcs, ctypes.c_int32, 4 * c.R).reshape(-1, 4)
        assert run_off[0] == 0 and run_off[-1] == c.R and (np.diff(run_off) >= 0).all()
        pi = np.zeros(N)
        for tk in fwd["tasks"]:
            pi[tk[0]] = _f32(tk[3])
        assert sorted(fwd["tasks"][:, 0].tolist()) == list(range(N)) == sorted(bwd["tasks"][:, 0].tolist())
        arc_recs = fwd["recs"][(fwd["recs"][:, 1] >= 0) & (fwd["recs"][:, 1] != 0)]
        assert len(arc_recs) == den.num_arcs                       # every arc exactly once ...
        assert (fwd["recs"][:, 1] < 0).sum() == c.R                # ... every run closed exactly once (no empty states here)
        sq = 3                                                     # one sequence of the minibatch
        e = np.exp(y[sq::B].astype(np.float64))                    # [T, P]
        alpha = np.zeros((2, N))
        alpha[0] = pi
        S = np.zeros(T + 1)
        S[0] = pi.sum()
        arun = np.zeros((T, c.R + 1))

        def walk(d, fn):
            for b in range(grid * bins):
                task = d["warp_task_off"][b]
                for k in range(d["warp_arc_off"][b], d["warp_arc_off"][b + 1]):
                    task = fn(d["recs"][k], task)

        for t in range(1, T + 1):
            prev, cur = alpha[(t - 1) & 1], alpha[t & 1]
            st = dict(racc=0.0, sacc=0.0)

            def frec(r, task):
                if r[1] >= 0:
                    st["racc"] += (prev[r[0] // rowbytes] + leaky * _f32(r[2]) * S[t - 1]) * _f32(r[1])
                else:
                    out = st["racc"] * e[t - 1, r[0] // rowbytes] / S[t - 1]
                    arun[t - 1, r[2] // rowbytes] = out
                    st["sacc"] += out
                    st["racc"] = 0.0
                    if r[1] & 1:
                        cur[fwd["tasks"][task][0]] = st["sacc"]
                        st["sacc"] = 0.0
                        return task + 1
                return task
            walk(fwd, frec)
            S[t] = cur.sum()
        tot = S[T] * (1.0 + leaky * pi.sum())
        logp = np.log(tot) + np.log(S[:T]).sum()
        beta = np.zeros((2, N))
        beta[T & 1] = 1.0 / tot
        gamma = np.zeros((T, P))
        for t in range(T - 1, -1, -1):
            nxt, cur = beta[(t + 1) & 1], beta[t & 1]
            bsum = leaky * (pi * nxt).sum()
            W = np.zeros(c.R + 1)
            for r in runs:                                          # phase A
                full = nxt[r[0] // rowbytes] + bsum
                gamma[t, r[1] // rowbytes] += arun[t, r[2] // rowbytes] * full
                W[r[2] // rowbytes] = full * e[t, r[1] // rowbytes]
            st = dict(X=0.0)

            def brec(r, task):                                      # phase B
                st["X"] += abs(_f32(r[2])) * W[r[0] // rowbytes]
                if r[2] < 0:
                    cur[bwd["tasks"][task][0]] = st["X"] / S[t]
                    st["X"] = 0.0
                    return task + 1
                return task
            walk(bwd, brec)
        import oracle_lib as ol
        oden = ol.OracleDen(den)
        lp_ref, d_ref, ok = ol.den_only(oden, B, T, y.astype(np.float64), leaky=leaky)
        gamma_ref = d_ref.reshape(T, B, P)[:, sq, :]
        assert np.abs(gamma.sum(axis=1) - 1.0).max() < 1e-6
        assert np.abs(gamma - gamma_ref).max() < 1e-6
        # the oracle's log-likelihood is the sum over all B sequences: check this sequence's share through its own run
        lp1, _, _ = ol.den_only(oden, 1, T, y[sq::B].astype(np.float64), leaky=leaky)
        assert lp1 == pytest.approx(logp, rel=1e-6)
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_untied_graphs_and_small_batches_keep_the_per_arc_kernels(mock):
    den_u = synth.make_den_fst(5, 150, 30, mean_deg=8.0)            # a pdf per arc: as many runs as arcs
    den_t = synth.make_den_fst_tied(5, 150, 30, mean_deg=8.0)
    for den, B, tied in ((den_u, 65, False), (den_t, 64, False), (den_t, 130, True)):
        sup = synth.make_supervision(6, den, B, 4, num_paths=2)
        y = synth.make_nnet_output(7, B, 4, 30).astype(np.float32)
        g, s = _handles(mock, den, sup)
        try:
            mock.mock_reset_launches()
            _compute(mock, g, s, y)
            names = [_kernel(mock.mock_launch_name(i).decode()) for i in range(mock.mock_launch_count())]
            den_names = [n for n in names if "den_" in n]
            if tied:   # 128 sequences run-based, the remaining 2 (padded to 32) per-arc
                assert den_names == ["den_tied_forward_kernel", "den_tied_backward_kernel", "den_forward_kernel",
                                     "den_backward_kernel"]
            else:
                assert den_names == ["den_forward_kernel", "den_backward_kernel"]
        finally:
            mock.my_lib_supervision_free(s)
            mock.my_lib_denominator_graph_free(g)


def test_tied_graph_too_large_for_shared_memory_stays_per_arc(mock, monkeypatch):
    """The run-based kernels keep a CTA's record streams in shared memory; a tied graph whose streams do not fit
    (2 M arcs: ~220 KB per CTA) is served by the per-arc kernels (which can read their arcs from global memory)."""
    monkeypatch.setenv("TORCHAIN_TIED", "1")
    den = synth.make_den_fst_tied(3, 12_000, 500, mean_deg=165.0)
    assert (den.num_arcs + 2 * den.num_states) / 148 * 16 > 200 * 1024
    B, T = 70, 2
    sup = synth.make_supervision(4, den, B, T, num_paths=1)
    y = synth.make_nnet_output(5, B, T, den.num_pdfs).astype(np.float32)
    g, s = _handles(mock, den, sup)
    try:
        assert mock.torchain_denominator_graph_is_tied(g) == 1
        mock.mock_reset_launches()
        _compute(mock, g, s, y)
        names = [_kernel(mock.mock_launch_name(i).decode()) for i in range(mock.mock_launch_count())]
        assert [n for n in names if "den_" in n] == ["den_forward_kernel", "den_backward_kernel"]
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


def test_host_entry_publishes_progress_only_at_chunk_boundaries(mock):
    """The backward launch releases the progress word at system scope (a ~1.7 us fence on a warp of CTA 0) only for the
    frame counts at which the host can send a chunk: DenParams::progress_mask carries exactly the values T - cb[k]."""
    den = synth.make_den_fst(13, 80, 16, mean_deg=3.0)
    B, T = 6, 50
    sup = synth.make_supervision(14, den, B, T, num_paths=2, weight=1.0)
    y = synth.make_nnet_output(15, B, T, den.num_pdfs).astype(np.float32)
    d = np.empty_like(y)
    res = np.zeros(3, dtype=np.float32)
    g, s = _handles(mock, den, sup)
    try:
        mock.mock_set_auto_progress(T)
        mock.mock_reset_launches()
        ok = mock.torchain_ComputeChainObjfAndDeriv_host(g, s, _p(y), y.shape[0], y.shape[1], y.shape[1], _p(res), _p(d),
                                                         y.shape[1], None, 0, 5e-5, 0.1, 0.0, 0)
        assert ok == 1, mock.torchain_last_error()
        c = _den_batches(mock)[1][1].c[0]               # the backward launch's parameters
        assert c.deriv_progress
        bounds = (ctypes.c_int * 64)()
        n = mock.torchain_host_chunk_bounds(T, bounds, 64)
        want = {T - bounds[k] for k in range(1, n - 1)}
        got = {v for v in range(512) if (c.progress_mask[v >> 6] >> (v & 63)) & 1}
        assert got == want and 0 < len(want) < T - 1
        # the device entry publishes nothing and leaves the mask at "every frame"
        mock.mock_reset_launches()
        _compute(mock, g, s, y)
        c2 = _den_batches(mock)[1][1].c[0]
        assert not c2.deriv_progress and all(c2.progress_mask[i] == 2**64 - 1 for i in range(8))
    finally:
        mock.mock_set_auto_progress(0)
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)


@pytest.mark.parametrize("B,subs", [(64, 2), (16, 4)])
def test_small_batch_streams_are_complete_and_aligned(mock, B, subs):
    """Batches of <= 64 sequences: every warp walks `subs` arc streams in lockstep, one per lane group.  The host
    puts states of equal arc count side by side and pads them to the same length, so that an end of state is
    warp-uniform (build_dir, "aligned streams").  Read the uploaded per-arc streams of both directions back and check
    (1) they hold every state once and every arc of the graph once, with its probability, pdf and the other state's
    initial probability, (2) the padding carries no probability, (3) the ends of state of a warp's streams coincide,
    (4) the padding is small."""
    P, N, T = 40, 3000, 3
    den = synth.make_den_fst(11, N, P, mean_deg=9.0)
    sup = synth.make_supervision(12, den, B, T, num_paths=2)
    y = synth.make_nnet_output(13, B, T, P).astype(np.float32)
    g, s = _handles(mock, den, sup)
    try:
        mock.mock_reset_launches()
        _compute(mock, g, s, y)
        names = [mock.mock_launch_name(i).decode() for i in range(mock.mock_launch_count())]
        i_f = [i for i, n in enumerate(names) if "den_forward_kernel" in n][0]
        c = DenBatch.from_buffer_copy(ctypes.string_at(mock.mock_launch_param0(i_f), ctypes.sizeof(DenBatch))).c[0]
        assert c.subs == subs and c.BP == 128 // subs
        grid, warps, rowbytes = 148, 14, 4 * c.BP
        prob = np.exp(-den.cost.astype(np.float32)).astype(np.float32)
        import oracle_lib as ol
        pi = ol.OracleDen(den).initial_probs().astype(np.float32)
        for d, (key, other) in enumerate([(den.dst, den.src), (den.src, den.dst)]):
            D = _dir_arrays(c.dir[d], grid * warps * subs, grid)
            recs, tasks, aoff, toff = D["recs"], D["tasks"], D["warp_arc_off"], D["warp_task_off"]
            assert sorted(tasks[:, 0].tolist()) == list(range(N))
            got, pad = [], 0
            for b in range(grid * warps * subs):
                task = toff[b]
                for k in range(aoff[b], aoff[b + 1]):
                    r = recs[k]
                    pr = abs(_f32(r[2]))
                    if pr != 0.0:
                        assert task < toff[b + 1]
                        got.append((int(tasks[task][0]), int(r[0]) // rowbytes, int(r[1]) // rowbytes,
                                    float(np.float32(pr)), float(np.float32(_f32(r[3])))))
                    else:
                        pad += 1
                    if r[2] < 0:
                        task += 1
                assert task == toff[b + 1]                       # every task of the stream is closed by an end flag
            want = sorted((int(key[a]), int(other[a]), int(den.ilabel[a]) - 1, float(prob[a]), float(pi[other[a]]))
                          for a in range(den.num_arcs))
            got.sort()
            assert [x[:3] for x in got] == [x[:3] for x in want]
            assert np.allclose([x[3:] for x in got], [x[3:] for x in want], rtol=1e-6, atol=0)   # expf of the library
            print("dir %d: %d padding records for %d arcs" % (d, pad, den.num_arcs))
            nstreams = int((np.diff(aoff) > 0).sum())
            assert pad <= 0.02 * den.num_arcs + 8 * nstreams       # alignment (neighbours in sorted order) + kStreamPad
            for w in range(grid * warps):
                ends = []
                for j in range(subs):
                    b = w * subs + j
                    z = recs[aoff[b]:aoff[b + 1], 2]
                    ends.append(np.nonzero(z < 0)[0].tolist())
                for j in range(1, subs):                          # a short last group leaves later streams shorter
                    assert ends[j] == ends[0][:len(ends[j])]
    finally:
        mock.my_lib_supervision_free(s)
        mock.my_lib_denominator_graph_free(g)
