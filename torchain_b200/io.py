"""Denominator-graph and supervision handles.

Mirrors the hot-path part of the reference's ``torchain/io.py``
(/root/reference/torchain/io.py:15-31,51-57): ``set_kaldi_device``,
``Supervision`` (opaque pointer with ``n_pdf/n_batch/n_frame/shape``, raises
``ValueError("null supervision ptr")`` on NULL) and ``DenominatorGraph(rspec,
n_pdf)``.  The Kaldi egs readers of that file (``Example``, ``RandExample``,
``open_example``) are outside the chain-loss path and are not provided; the two
handles additionally offer array constructors (``from_arrays``) so a caller can
hand over an FST it already holds in memory.
"""
import ctypes

import numpy as np

from . import _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_kaldi_device(device_id=0):
    """Reference: io.py:15-17.  No global device state exists here; validates the index."""
    lib = _lib.load()
    if lib.my_lib_set_kaldi_device(int(device_id)) != 1:
        raise RuntimeError(_lib.last_error())


class Supervision:
    """Reference: io.py:20-31.  Owns a ``kaldi::chain::Supervision``-equivalent handle."""

    def __init__(self, ptr):
        self.ptr = ptr
        if not self.ptr:  # ffi.NULL in the reference
            raise ValueError("null supervision ptr")
        lib = _lib.load()
        self.n_pdf = lib.my_lib_supervision_num_pdf(self.ptr)
        self.n_batch = lib.my_lib_supervision_num_sequence(self.ptr)
        self.n_frame = lib.my_lib_supervision_num_frame(self.ptr)
        self.shape = (self.n_batch, self.n_frame, self.n_pdf)

    @classmethod
    def from_arrays(cls, weight, num_sequences, frames_per_sequence, label_dim,
                    arc_begin, arc_dst, arc_label, arc_cost, final_cost):
        """Build from the merged supervision FST (states in time order, state 0 = start,
        ``arc_label = pdf_id + 1``, tropical costs, ``final_cost = inf`` for non-final)."""
        lib = _lib.load()
        arc_begin = np.ascontiguousarray(arc_begin, dtype=np.int32)
        arc_dst = np.ascontiguousarray(arc_dst, dtype=np.int32)
        arc_label = np.ascontiguousarray(arc_label, dtype=np.int32)
        arc_cost = np.ascontiguousarray(arc_cost, dtype=np.float32)
        final_cost = np.ascontiguousarray(final_cost, dtype=np.float32)
        num_states = int(final_cost.shape[0])
        if arc_begin.shape[0] != num_states + 1:
            raise ValueError("arc_begin must have num_states + 1 entries")
        n_arcs = int(arc_begin[-1])
        if not (arc_dst.shape[0] == arc_label.shape[0] == arc_cost.shape[0] == n_arcs):
            raise ValueError("arc arrays must have arc_begin[-1] entries")
        ptr = lib.torchain_supervision_from_arrays(
            float(weight), int(num_sequences), int(frames_per_sequence), int(label_dim), num_states,
            _ptr(arc_begin), _ptr(arc_dst), _ptr(arc_label), _ptr(arc_cost), _ptr(final_cost))
        if not ptr:
            raise ValueError("invalid supervision: " + _lib.last_error())
        return cls(ptr)

    @classmethod
    def from_kaldi_bytes(cls, data):
        """Parse one binary Kaldi ``Supervision`` object (bytes starting at ``<Supervision>``).
        Returns ``(supervision, bytes_consumed)``."""
        lib = _lib.load()
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
        consumed = ctypes.c_int64(0)
        ptr = lib.torchain_supervision_from_kaldi_bytes(ctypes.cast(buf, ctypes.c_void_p), len(data),
                                                        ctypes.byref(consumed))
        if not ptr:
            raise ValueError("invalid supervision: " + _lib.last_error())
        return cls(ptr), int(consumed.value)

    @classmethod
    def from_chain_egs(cls, data, max_examples=None):
        """The supervisions of the binary nnet3 chain examples in `data` (bytes of an egs ark, e.g. what
        ``nnet3-chain-copy-egs ... ark:-`` writes), one handle per example, found by scanning for the
        ``<Nnet3ChainEg>`` token like the reference's indexer (/root/reference/example/chime5/ark2scp.py:24-40).
        What the reference's readers hand out per example (``Example.supervision``, io.py:74-76) -- without the
        features.  ``Supervision.merge(handles)`` makes the minibatch."""
        lib = _lib.load()
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
        base = ctypes.addressof(buf)
        out, pos = [], 0
        while pos < len(data) and (max_examples is None or len(out) < max_examples):
            if data.find(b"<Nnet3ChainEg>", pos) < 0:
                break
            consumed = ctypes.c_int64(0)
            ptr = lib.torchain_supervision_from_chain_egs(ctypes.c_void_p(base + pos), len(data) - pos,
                                                          ctypes.byref(consumed))
            if not ptr:
                raise ValueError("invalid chain example: " + _lib.last_error())
            out.append(cls(ptr))
            pos += int(consumed.value)
        if not out:
            raise ValueError("invalid chain example: no <Nnet3ChainEg> token in the buffer")
        return out

    @classmethod
    def merge(cls, supervisions):
        """One minibatch supervision from a list of examples, appended sequence-major: what the
        reference's egs reader obtains from Kaldi's MergeChainExamples
        (/root/reference/src/my_lib_example_rand.cpp:160).  Rows of the matching nnet_output are
        ``t * B_total + (sequences of example 0, then of example 1, ...)``."""
        sups = list(supervisions)
        if not sups:
            raise ValueError("invalid supervision: empty list")
        arr = (ctypes.c_void_p * len(sups))(*[s.ptr for s in sups])
        ptr = _lib.load().torchain_supervision_merge(arr, len(sups))
        if not ptr:
            raise ValueError("invalid supervision: " + _lib.last_error())
        return cls(ptr)

    def prefetch(self, device=None, stream=None):
        """Start the asynchronous upload of the device image on `stream` (default: the current stream) of
        `device`, e.g. from a data-loading thread while the previous minibatch is still being computed.
        Optional -- the first chain_loss on a device uploads it otherwise.  Never synchronises the device."""
        import torch
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device("cuda", device).index if not isinstance(device, int) else device
        if stream is None:
            stream = torch.cuda.current_stream(device)
        if _lib.load().torchain_supervision_prefetch(self.ptr, int(device), ctypes.c_void_p(stream.cuda_stream)) != 1:
            raise RuntimeError("supervision prefetch failed: " + _lib.last_error())
        return self

    @property
    def num_chains(self):
        return _lib.load().torchain_supervision_num_chains(self.ptr)

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                _lib.load().my_lib_supervision_free(ptr)
            except Exception:  # interpreter shutdown
                pass


class DenominatorGraph:
    """Reference: io.py:51-57.  ``DenominatorGraph(rspec, n_pdf)`` reads den.fst."""

    def __init__(self, rspec, n_pdf, _ptr_=None):
        self.rspec = rspec
        if _ptr_ is not None:
            self.ptr = _ptr_
        else:
            self.ptr = _lib.load().my_lib_denominator_graph_new(str(rspec).encode(), int(n_pdf))
        if not self.ptr:
            raise ValueError("cannot build denominator graph: " + _lib.last_error())

    @classmethod
    def from_arrays(cls, num_states, src, dst, ilabel, cost, start, final_cost, n_pdf):
        lib = _lib.load()
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        ilabel = np.ascontiguousarray(ilabel, dtype=np.int32)
        cost = np.ascontiguousarray(cost, dtype=np.float32)
        final_cost = np.ascontiguousarray(final_cost, dtype=np.float32)
        if final_cost.shape[0] != num_states:
            raise ValueError("final_cost must have num_states entries")
        if not (src.shape[0] == dst.shape[0] == ilabel.shape[0] == cost.shape[0]):
            raise ValueError("arc arrays must have equal length")
        ptr = lib.torchain_denominator_graph_from_arrays(
            int(num_states), int(src.shape[0]), _ptr(src), _ptr(dst), _ptr(ilabel), _ptr(cost),
            int(start), _ptr(final_cost), int(n_pdf))
        if not ptr:
            raise ValueError("invalid denominator graph: " + _lib.last_error())
        return cls("<arrays>", n_pdf, _ptr_=ptr)

    @property
    def n_state(self):
        return _lib.load().torchain_denominator_graph_num_states(self.ptr)

    @property
    def n_arc(self):
        return _lib.load().torchain_denominator_graph_num_arcs(self.ptr)

    @property
    def n_pdf(self):
        return _lib.load().torchain_denominator_graph_num_pdfs(self.ptr)

    @property
    def n_run(self):
        """Distinct (destination state, pdf) pairs; ~2 per state in a real den.fst."""
        return _lib.load().torchain_denominator_graph_num_runs(self.ptr)

    @property
    def tied(self):
        """True if the run-based (tied-pdf) kernels serve this graph, False for the per-arc kernels."""
        return bool(_lib.load().torchain_denominator_graph_is_tied(self.ptr))

    @property
    def initial_probs(self):
        out = np.empty(self.n_state, dtype=np.float32)
        if _lib.load().torchain_denominator_graph_initial_probs(self.ptr, _ptr(out)) != 1:
            raise RuntimeError(_lib.last_error())
        return out

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                _lib.load().my_lib_denominator_graph_free(ptr)
            except Exception:
                pass
