"""CPU suite, part 2: the C-ABI library loads without a GPU, exports every symbol
include/torchain_b200.h declares, and its host-side handle logic (no compute
calls) behaves like the reference's handle layer (/root/reference/torchain/io.py:20-31,51-57)."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

import oracle_lib as ol
from torchain_b200 import _lib, io, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "torchain_b200.h")
HEADERS = [HEADER, os.path.join(ROOT, "include", "torchain_b200_profile.h")]


def declared_symbols():
    names = set()
    for h in HEADERS:
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b((?:my_lib|torchain)_\w+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.torchain_abi_version() == 1
    assert lib.my_lib_supervision_new(None) is None
    assert "egs readers" in _lib.last_error()


def test_den_graph_from_arrays_matches_oracle_initial_probs():
    den = synth.make_den_fst(7, 300, 40, mean_deg=4.0)
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    assert (g.n_state, g.n_arc, g.n_pdf) == (den.num_states, den.num_arcs, den.num_pdfs)
    ref = ol.OracleDen(den).initial_probs()
    assert np.abs(g.initial_probs - ref).max() < 1e-7


def test_den_graph_rejects_bad_arrays():
    den = synth.make_den_fst(7, 20, 5, mean_deg=2.0)
    bad = den.ilabel.copy()
    bad[0] = 0   # label 0 would be pdf -1
    with pytest.raises(ValueError):
        io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, bad, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    with pytest.raises(ValueError):
        io.DenominatorGraph("/nonexistent/den.fst", 10)


def write_openfst_binary(path, den):
    """OpenFst VectorFst<StdArc> binary layout."""
    order = np.argsort(den.src, kind="stable")
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 2125659606))
        for s in (b"vector", b"standard"):
            f.write(struct.pack("<i", len(s)) + s)
        f.write(struct.pack("<iiQqqq", 2, 0, 0, den.start, den.num_states, den.num_arcs))
        k = 0
        counts = np.bincount(den.src, minlength=den.num_states)
        for s in range(den.num_states):
            f.write(struct.pack("<fq", float(den.final_cost[s]), int(counts[s])))
            for _ in range(counts[s]):
                a = order[k]
                k += 1
                f.write(struct.pack("<iifi", int(den.ilabel[a]), int(den.ilabel[a]), float(den.cost[a]),
                                    int(den.dst[a])))


def test_den_graph_from_openfst_binary_and_text(tmp_path):
    den = synth.make_den_fst(11, 60, 9, mean_deg=3.0)
    ref = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                          den.final_cost, den.num_pdfs).initial_probs
    p = str(tmp_path / "den.fst")
    write_openfst_binary(p, den)
    g = io.DenominatorGraph(p, den.num_pdfs)
    assert (g.n_state, g.n_arc) == (den.num_states, den.num_arcs)
    assert np.array_equal(g.initial_probs, ref)
    # AT&T text form
    t = str(tmp_path / "den.txt")
    with open(t, "w") as f:
        for a in range(den.num_arcs):
            f.write("%d %d %d %d %.9g\n" % (den.src[a], den.dst[a], den.ilabel[a], den.ilabel[a], den.cost[a]))
        for s in range(den.num_states):
            f.write("%d %.9g\n" % (s, den.final_cost[s]))
    g2 = io.DenominatorGraph(t, den.num_pdfs)
    assert (g2.n_state, g2.n_arc) == (den.num_states, den.num_arcs)
    assert np.abs(g2.initial_probs - ref).max() < 1e-6


def test_supervision_handle_shape_and_split():
    den = synth.make_den_fst(3, 50, 12, mean_deg=3.0)
    for bc in (False, True):
        sup = synth.make_supervision(4, den, 5, 7, num_paths=3, weight=0.5, boundary_final_costs=bc)
        h = io.Supervision.from_arrays(sup.weight, sup.num_sequences, sup.frames_per_sequence, sup.label_dim,
                                       sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)
        assert h.shape == (5, 7, 12) and (h.n_batch, h.n_frame, h.n_pdf) == (5, 7, 12)
        assert h.num_chains == 5   # the merged FST factorises per sequence


def test_supervision_unsplittable_falls_back_to_one_chain():
    den = synth.make_den_fst(3, 50, 12, mean_deg=3.0)
    sup = synth.make_supervision(4, den, 3, 6, num_paths=3)
    # perturb ONE copied start arc at the first sequence boundary so the coupling is no longer rank one
    cost = sup.arc_cost.copy()
    times = np.full(sup.num_states, -1)
    times[0] = 0
    for s in range(sup.num_states):
        for a in range(sup.arc_begin[s], sup.arc_begin[s + 1]):
            times[sup.arc_dst[a]] = times[s] + 1
    boundary = np.where(times == 6)[0]
    if len(boundary) < 2:
        pytest.skip("needs two boundary states")
    cost[sup.arc_begin[boundary[1]]] += 0.5
    h = io.Supervision.from_arrays(sup.weight, 3, 6, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                                   cost, sup.final_cost)
    assert h.num_chains == 1


def test_supervision_null_and_invalid():
    with pytest.raises(ValueError, match="null supervision ptr"):   # reference io.py:23-24
        io.Supervision(None)
    den = synth.make_den_fst(3, 30, 6, mean_deg=2.0)
    sup = synth.make_supervision(4, den, 2, 4, num_paths=2)
    with pytest.raises(ValueError):   # wrong frames_per_sequence
        io.Supervision.from_arrays(1.0, 2, 5, sup.label_dim, sup.arc_begin, sup.arc_dst, sup.arc_label,
                                   sup.arc_cost, sup.final_cost)


def test_chain_results_formula():
    """Reference functions.py:9-19."""
    from torchain_b200.functions import ChainResults
    r = ChainResults()
    r.data[0], r.data[1], r.data[2] = -120.0, -3.0, 60.0
    assert float(r.loss) == 2.0
    assert "ChainResults(loss=2.0" in repr(r)


def test_to2d_row_order():
    """Reference functions.py:118-125: (B, C, T) -> row t*B + b."""
    import torch
    from torchain_b200.functions import to2d
    x = torch.arange(2 * 3 * 4, dtype=torch.float32).view(2, 3, 4)
    y = to2d(x)
    assert y.shape == (8, 3)
    for b in range(2):
        for t in range(4):
            assert torch.equal(y[t * 2 + b], x[b, :, t])


def test_chain_loss_requires_cuda():
    """Reference functions.py:67: CUDA-only; there is no CPU fallback."""
    import torch
    from torchain_b200.functions import chain_loss
    with pytest.raises(AssertionError, match="Only CUDA"):
        chain_loss(torch.zeros(4, 3), None, None)


def kaldi_supervision_bytes(sup, end2end_token=False):
    """Byte layout of chain::Supervision::Write(os, binary=true) (Kaldi) with the FST written by
    fst::StdCompactAcceptorFst::WriteFst (OpenFst): see the parser's header comment."""
    out = b"<Supervision> <Weight> " + struct.pack("<Bf", 4, sup.weight)
    out += b"<NumSequences> " + struct.pack("<Bi", 4, sup.num_sequences)
    out += b"<FramesPerSeq> " + struct.pack("<Bi", 4, sup.frames_per_sequence)
    out += b"<LabelDim> " + struct.pack("<Bi", 4, sup.label_dim)
    if end2end_token:
        out += b"<End2End> F"
    S = sup.num_states
    offs, elems = [0], b""
    n = 0
    for s in range(S):
        if np.isfinite(sup.final_cost[s]):
            elems += struct.pack("<ifi", -1, float(sup.final_cost[s]), -1)
            n += 1
        for a in range(sup.arc_begin[s], sup.arc_begin[s + 1]):
            elems += struct.pack("<ifi", int(sup.arc_label[a]), float(sup.arc_cost[a]), int(sup.arc_dst[a]))
            n += 1
        offs.append(n)
    out += struct.pack("<i", 2125659606)
    for t in (b"compact_acceptor", b"standard"):
        out += struct.pack("<i", len(t)) + t
    out += struct.pack("<iiQqqq", 2, 0, 0, 0, S, sup.num_arcs)
    out += struct.pack("<%dI" % (S + 1), *offs) + elems + b"</Supervision> "
    return out


def test_supervision_from_kaldi_bytes_round_trip():
    den = synth.make_den_fst(3, 40, 9, mean_deg=3.0)
    sup = synth.make_supervision(4, den, 4, 5, num_paths=3, weight=0.5, boundary_final_costs=True)
    for e2e in (False, True):
        blob = kaldi_supervision_bytes(sup, end2end_token=e2e)
        h, used = io.Supervision.from_kaldi_bytes(blob + b"trailing")
        assert used == len(blob) - 1 or used == len(blob)   # the token's trailing space is consumed
        assert h.shape == (4, 5, 9) and h.num_chains == 4
    with pytest.raises(ValueError):
        io.Supervision.from_kaldi_bytes(blob[:40])
    with pytest.raises(ValueError):
        io.Supervision.from_kaldi_bytes(b"<Nnet3ChainEg> ")


def _sup_handle(sup):
    return io.Supervision.from_arrays(sup.weight, sup.num_sequences, sup.frames_per_sequence, sup.label_dim,
                                      sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)


def test_supervision_merge_host_logic():
    """Minibatch assembly (Kaldi MergeChainExamples, reference src/my_lib_example_rand.cpp:160):
    sizes add up, chains stay per sequence, inconsistent examples are refused like Kaldi does."""
    den = synth.make_den_fst(3, 40, 10, mean_deg=3.0)
    parts = [synth.make_supervision(10 + i, den, b, 5, num_paths=2) for i, b in enumerate((1, 2, 1, 3))]
    hs = [_sup_handle(p) for p in parts]
    m = io.Supervision.merge(hs)
    assert m.shape == (7, 5, 10)
    assert m.num_chains == sum(h.num_chains for h in hs) == 7
    assert io.Supervision.merge([hs[1]]).shape == hs[1].shape
    # the inputs stay usable (the merge copies)
    del m
    assert hs[0].shape == (1, 5, 10)
    with pytest.raises(ValueError, match="empty"):
        io.Supervision.merge([])
    other_T = _sup_handle(synth.make_supervision(20, den, 1, 6, num_paths=2))
    with pytest.raises(ValueError, match="inconsistent"):
        io.Supervision.merge([hs[0], other_T])
    other_w = _sup_handle(synth.make_supervision(21, den, 1, 5, num_paths=2, weight=0.5))
    with pytest.raises(ValueError, match="inconsistent"):
        io.Supervision.merge([hs[0], other_w])


@pytest.mark.parametrize("seed", [1, 2])
def test_binary_parsers_reject_corrupt_input_without_crashing(seed, tmp_path):
    """The den.fst reader and the Kaldi Supervision parser read files the library does not control:
    truncations, flipped bytes and absurd counts must end in an error (or a still-valid object), never in a
    crash or an out-of-memory attempt.  (6000 mutations per parser were run once by hand; this keeps 400.)"""
    import random
    rng = random.Random(seed)
    den = synth.make_den_fst(3, 40, 10, mean_deg=3.0)
    sup = synth.make_supervision(4, den, 3, 5, num_paths=2)

    def mutate(raw, width):
        b = bytearray(raw)
        mode = rng.randrange(3)
        if mode == 0:
            return bytes(b[:rng.randrange(len(b))])
        if mode == 1:
            for _ in range(rng.randrange(1, 6)):
                b[rng.randrange(len(b))] = rng.randrange(256)
            return bytes(b)
        i = rng.randrange(len(b) - width)
        big = rng.choice([0, -1, 2 ** (8 * width - 2), -2 ** (8 * width - 2), 10 ** 6])
        b[i:i + width] = int(big).to_bytes(width, "little", signed=True)
        return bytes(b)

    blob = kaldi_supervision_bytes(sup)
    accepted = 0
    for _ in range(200):
        try:
            h, _used = io.Supervision.from_kaldi_bytes(mutate(blob, 4))
            assert h.shape[2] >= 1
            accepted += 1
        except ValueError:
            pass
    path = str(tmp_path / "den.fst")
    write_openfst_binary(path, den)
    raw = open(path, "rb").read()
    bad = str(tmp_path / "bad.fst")
    for _ in range(200):
        with open(bad, "wb") as f:
            f.write(mutate(raw, 8))
        try:
            g = io.DenominatorGraph(bad, den.num_pdfs)
            assert g.n_pdf == den.num_pdfs
            accepted += 1
        except (ValueError, RuntimeError, OSError):
            pass
    assert accepted < 400   # most mutations must be caught


@pytest.mark.parametrize("seed", [1, 2])
def test_array_constructors_reject_corrupt_input_without_crashing(seed):
    """Out-of-range indices, broken CSR offsets (regression: arc_begin[0] = INT_MIN used to be dereferenced),
    absurd sizes and non-finite costs in the array constructors: an error or a valid handle, never a crash."""
    rng = np.random.default_rng(seed)
    den = synth.make_den_fst(3, 40, 10, mean_deg=3.0)
    sup = synth.make_supervision(4, den, 3, 5, num_paths=2)
    ints = [-1, -2 ** 31, 2 ** 31 - 1, 0, 10 ** 6, 7, den.num_states]
    for _ in range(300):
        d = dict(N=den.num_states, src=den.src.copy(), dst=den.dst.copy(), il=den.ilabel.copy(), cost=den.cost.copy(),
                 start=den.start, P=den.num_pdfs)
        k, i, v = rng.integers(0, 7), rng.integers(0, den.src.shape[0]), int(rng.choice(ints))
        if k == 0: d["src"][i] = v
        elif k == 1: d["dst"][i] = v
        elif k == 2: d["il"][i] = v
        elif k == 3: d["cost"][i] = rng.choice([np.nan, np.inf, -np.inf, 1e30, -1e30])
        elif k == 4: d["start"] = v
        elif k == 5: d["N"] = int(rng.choice([0, -1, 1, 10 ** 6, den.num_states - 1]))
        else: d["P"] = int(rng.choice([0, -1, 1, 5]))
        try:
            io.DenominatorGraph.from_arrays(d["N"], d["src"], d["dst"], d["il"], d["cost"], d["start"],
                                            den.final_cost, d["P"])
        except (ValueError, RuntimeError, AssertionError):
            pass
    for _ in range(300):
        ab, ad, al = sup.arc_begin.copy(), sup.arc_dst.copy(), sup.arc_label.copy()
        ac, fc = sup.arc_cost.copy(), sup.final_cost.copy()
        B, T = sup.num_sequences, sup.frames_per_sequence
        k, v = rng.integers(0, 7), int(rng.choice(ints))
        if k == 0: ab[rng.integers(0, ab.shape[0])] = v
        elif k == 1: ad[rng.integers(0, ad.shape[0])] = v
        elif k == 2: al[rng.integers(0, al.shape[0])] = v
        elif k == 3: ac[rng.integers(0, ac.shape[0])] = rng.choice([np.nan, np.inf, -np.inf])
        elif k == 4: fc[rng.integers(0, fc.shape[0])] = rng.choice([np.nan, 0.0, -np.inf])
        elif k == 5: B = int(rng.choice([0, -1, 1, 4, 10 ** 6]))
        else: T = int(rng.choice([0, -1, 1, 4, 10 ** 6]))
        try:
            io.Supervision.from_arrays(sup.weight, B, T, sup.label_dim, ab, ad, al, ac, fc)
        except (ValueError, RuntimeError, AssertionError):
            pass
    with pytest.raises(ValueError, match="arc_begin"):
        ab = sup.arc_begin.copy()
        ab[0] = -2 ** 31
        io.Supervision.from_arrays(sup.weight, sup.num_sequences, sup.frames_per_sequence, sup.label_dim, ab,
                                   sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)


def test_host_chunk_bounds_schedule():
    """Chunk boundaries of the host-buffer entry: a partition of [0, T) into whole-frame chunks,
    about 16 of them, single frames at both ends once there are more than 16 frames."""
    import ctypes
    lib = _lib.load()
    assert lib.torchain_host_chunk_bounds(0, None, 0) == 0
    for T in list(range(1, 300)) + [500, 1000, 1599, 5000]:
        n = lib.torchain_host_chunk_bounds(T, None, 0)
        buf = (ctypes.c_int * n)()
        assert lib.torchain_host_chunk_bounds(T, buf, n) == n
        b = list(buf)
        assert b[0] == 0 and b[-1] == T
        sizes = [b[i + 1] - b[i] for i in range(n - 1)]
        assert all(sz >= 1 for sz in sizes)
        assert len(sizes) <= 16 + 2 * 12
        base = max(1, (T + 15) // 16)
        assert max(sizes) <= base
        if T > 32:
            assert sizes[0] == 1 and sizes[-1] == 1
            assert sizes == sizes[::-1] or sizes[:3] == sizes[::-1][:3]


def test_text_fst_and_dead_end_graphs_end_in_errors(tmp_path):
    """Hostile text FSTs (negative or huge state ids, a size field that cannot be allocated) and a graph
    whose probability mass dies out end in ValueError with a message -- no out-of-bounds write, no
    exception through the C ABI, no NaN initial probabilities."""
    cases = {
        "neg_final.txt": "0 1 1 1 0.5\n1 0 1 1 0.5\n-7 0.0\n",
        "neg_arc.txt": "0 -3 1 1 0.5\n0 0.0\n",
        "huge_state.txt": "0 2000000000 1 1 0.5\n0 0.0\n",
        "huge_final.txt": "0 0 1 1 0.5\n2147483647 0.0\n",
    }
    for name, text in cases.items():
        p = str(tmp_path / name)
        with open(p, "w") as f:
            f.write(text)
        with pytest.raises(ValueError):
            io.DenominatorGraph(p, 4)
        assert _lib.last_error() != ""
    # binary header announcing 2^30 states in a 100-byte file
    p = str(tmp_path / "liar.fst")
    with open(p, "wb") as f:
        f.write(struct.pack("<i", 2125659606))
        for s in (b"vector", b"standard"):
            f.write(struct.pack("<i", len(s)) + s)
        f.write(struct.pack("<iiQqqq", 2, 0, 0, 0, 1 << 30, 5))
        f.write(b"\0" * 20)
    with pytest.raises(ValueError):
        io.DenominatorGraph(p, 4)
    # start state -> dead end: Kaldi's SetInitialProbs would divide by zero
    with pytest.raises(ValueError, match="no outgoing arcs"):
        io.DenominatorGraph.from_arrays(2, np.array([0], dtype=np.int32), np.array([1], dtype=np.int32),
                                        np.array([1], dtype=np.int32), np.array([0.1], dtype=np.float32), 0,
                                        np.array([np.inf, 0.0], dtype=np.float32), 1)


def test_header_is_plain_c_and_links_from_a_c_host(tmp_path):
    """The boundary is a C ABI: include/torchain_b200.h must compile as C99 (no C++-isms, no torch or CUDA
    types) and a C host must be able to link the library and call it -- here the device-free entry points."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "torchain_b200.h"
int main(void) {
  int b[64];
  int n = torchain_host_chunk_bounds(50, b, 64);
  if (torchain_abi_version() < 1 || n < 2 || b[0] != 0 || b[n - 1] != 50) return 1;
  /* error convention: NULL / 0 plus a message, never an abort */
  if (my_lib_denominator_graph_new("/nonexistent/den.fst", 10) != NULL) return 2;
  if (strlen(torchain_last_error()) == 0) return 3;
  if (my_lib_supervision_num_pdf(NULL) != 0) return 4;
  printf("ok %d\n", n);
  return 0;
}
''')
    exe = str(tmp_path / "host")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.dirname(HEADER), str(src),
                    "-o", exe, "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.startswith("ok ")


@pytest.mark.parametrize("seed", [1, 2])
def test_text_fst_reader_rejects_corrupt_input_without_crashing(seed, tmp_path):
    """Same for the AT&T text form of den.fst: tokens replaced by negative / huge / non-numeric ones, lines cut."""
    import random
    rng = random.Random(seed)
    den = synth.make_den_fst(3, 40, 10, mean_deg=3.0)
    lines = ["%d %d %d %d %.9g" % (den.src[a], den.dst[a], den.ilabel[a], den.ilabel[a], den.cost[a])
             for a in range(den.num_arcs)] + ["%d %.9g" % (s, den.final_cost[s]) for s in range(den.num_states)]
    junk = ["-1", "-2147483648", "2147483647", "99999999999999999999", "1e400", "nan", "inf", "-inf", "x", "", "0x10", "3.7"]
    path = str(tmp_path / "den.txt")
    accepted = 0
    for _ in range(150):
        ls = list(lines)
        for _ in range(rng.randrange(1, 4)):
            i = rng.randrange(len(ls))
            tok = ls[i].split()
            mode = rng.randrange(3)
            if mode == 0 and tok:
                tok[rng.randrange(len(tok))] = rng.choice(junk)
            elif mode == 1 and tok:
                tok = tok[:rng.randrange(len(tok))]
            else:
                tok = tok + [rng.choice(junk)]
            ls[i] = " ".join(tok)
        with open(path, "w") as f:
            f.write("\n".join(ls) + "\n")
        try:
            g = io.DenominatorGraph(path, den.num_pdfs)
            assert g.n_pdf == den.num_pdfs and np.isfinite(g.initial_probs).all()
            accepted += 1
        except (ValueError, RuntimeError, OSError):
            pass
    assert accepted < 150


def chain_egs_bytes(sups, rng, keys=None):
    """A binary egs ark as nnet3-chain-get-egs writes it, reduced to what the scanner looks at: per example
    ``key \\0B<Nnet3ChainEg> ...inputs... <NnetChainSup> output ...indexes... <Supervision>...</Supervision> <DW2>
    ...weights... </NnetChainSup> </Nnet3ChainEg>`` with opaque bytes where features / indexes / weights would be."""
    out = b""
    for k, sup in enumerate(sups):
        junk = lambda n: bytes(rng.integers(0, 256, size=n, dtype=np.uint8)).replace(b"<", b"(")
        out += (keys[k] if keys else b"utt%d-0" % k) + b" \x00B<Nnet3ChainEg> <NumInputs> " + struct.pack("<Bi", 4, 2)
        out += b"<NnetIo> input " + junk(400) + b"</NnetIo> <NnetIo> ivector " + junk(60) + b"</NnetIo> "
        out += b"<NumOutputs> " + struct.pack("<Bi", 4, 1) + b"<NnetChainSup> output <I1V> " + junk(90)
        out += kaldi_supervision_bytes(sup) + b"<DW2> " + junk(70) + b"</NnetChainSup> </Nnet3ChainEg> "
    return out


def test_supervision_from_chain_egs_scans_an_ark():
    den = synth.make_den_fst(3, 40, 9, mean_deg=3.0)
    rng = np.random.default_rng(5)
    sups = [synth.make_supervision(10 + k, den, 1 + k % 2, 6, num_paths=2, weight=1.0) for k in range(4)]
    ark = chain_egs_bytes(sups, rng)
    hs = io.Supervision.from_chain_egs(ark)
    assert [h.shape for h in hs] == [(sp.num_sequences, 6, 9) for sp in sups]
    assert len(io.Supervision.from_chain_egs(ark, max_examples=2)) == 2
    mb = io.Supervision.merge(hs)
    assert mb.shape == (sum(sp.num_sequences for sp in sups), 6, 9) and mb.num_chains == mb.n_batch
    # a bare Supervision object announced by Kaldi's binary marker parses too
    h, used = io.Supervision.from_kaldi_bytes(b"\x00B" + kaldi_supervision_bytes(sups[0]))
    assert h.shape == hs[0].shape
    with pytest.raises(ValueError, match="text-mode"):
        io.Supervision.from_chain_egs(b"utt1 <Nnet3ChainEg> <NumInputs> 2 ...")
    with pytest.raises(ValueError, match="without a <Supervision>"):
        io.Supervision.from_chain_egs(b"utt1 \x00B<Nnet3ChainEg> <NumInputs> </Nnet3ChainEg> ")
    with pytest.raises(ValueError):
        io.Supervision.from_chain_egs(b"no examples here")
    with pytest.raises(ValueError):   # truncated inside the supervision
        io.Supervision.from_chain_egs(ark[:ark.index(b"<Supervision>") + 60])


def test_vectorised_openfst_writer_equals_the_per_arc_writer(tmp_path):
    den = synth.make_den_fst(12, 200, 30, mean_deg=5.0)
    a, b = str(tmp_path / "a.fst"), str(tmp_path / "b.fst")
    write_openfst_binary(a, den)
    synth.write_openfst_binary(b, den)
    assert open(a, "rb").read() == open(b, "rb").read()
