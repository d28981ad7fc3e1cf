"""Synthetic inputs for the chain loss: denominator graphs, supervisions, nnet outputs.

The reference has no runnable fixtures for this path (its tests need private
CHiME-5 egs, /root/reference/test/test.py:36-38,83-86, or Kaldi's random
TransitionModel machinery, /root/reference/src/my_lib_chain.cpp:138-213), so the
configurations of BASELINE.json are generated here:

* den graph: N states, out-degree ~ Poisson(mean_deg) plus a self-loop and a
  ring arc (keeps the graph strongly connected), pdfs uniform, probabilities
  Dirichlet-normalised per source state and scaled by (1 - final_prob).
* supervision for each sequence: the union of K random T-step walks through the
  den graph, merged into a lattice keyed by (frame, den state); arc label =
  that den arc's pdf + 1, arc cost = -log(transition prob); the first arc also
  carries -log(initial prob of the walk's first state).  This is what composing
  a supervision with Kaldi's normalization FST produces
  (/root/reference/src/my_lib_chain.cpp:200-204) and makes ``objf <= 0`` hold
  (/root/reference/src/chain-supervision-test.hpp:285).  The per-sequence FSTs
  are then concatenated the way Kaldi's MergeChainExamples/AppendSupervision
  does (Concat + RmEpsilon): the final states of sequence k-1 carry copies of
  sequence k's start arcs.

Pure numpy; no dependency on the CUDA library or on oracle/.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class DenFst:
    num_states: int
    num_pdfs: int
    start: int
    src: np.ndarray      # int32 [A]
    dst: np.ndarray      # int32 [A]
    ilabel: np.ndarray   # int32 [A]  (pdf + 1)
    cost: np.ndarray     # float32 [A]
    final_cost: np.ndarray  # float32 [N], inf = non-final

    @property
    def num_arcs(self):
        return int(self.src.shape[0])


@dataclass
class SupFst:
    weight: float
    num_sequences: int
    frames_per_sequence: int
    label_dim: int
    arc_begin: np.ndarray   # int32 [S+1]
    arc_dst: np.ndarray     # int32 [A]
    arc_label: np.ndarray   # int32 [A]
    arc_cost: np.ndarray    # float32 [A]
    final_cost: np.ndarray  # float32 [S]

    @property
    def num_states(self):
        return int(self.final_cost.shape[0])

    @property
    def num_arcs(self):
        return int(self.arc_dst.shape[0])


def make_den_fst(seed, num_states, num_pdfs, mean_deg=28.0, final_prob=0.05, max_deg=None):
    """Random strongly-connected HMM graph (arcs grouped by source state)."""
    rng = np.random.default_rng(seed)
    N = int(num_states)
    extra = rng.poisson(mean_deg, size=N).astype(np.int64)
    if max_deg is not None:
        extra = np.minimum(extra, max_deg)
    deg = extra + 2  # + self-loop + ring arc
    A = int(deg.sum())
    begin = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(deg, out=begin[1:])
    src = np.repeat(np.arange(N, dtype=np.int64), deg)
    dst = rng.integers(0, N, size=A, dtype=np.int64)
    dst[begin[:-1]] = np.arange(N)                 # self-loop
    dst[begin[:-1] + 1] = (np.arange(N) + 1) % N   # ring
    pdf = rng.integers(0, num_pdfs, size=A, dtype=np.int64)
    g = rng.gamma(1.0, size=A)                     # Dirichlet(1) per source after normalisation
    tot = np.add.reduceat(g, begin[:-1])
    prob = g / np.repeat(tot, deg) * (1.0 - final_prob)
    prob = np.maximum(prob, 1e-6)
    cost = (-np.log(prob)).astype(np.float32)
    final_cost = np.full(N, -np.log(final_prob) if final_prob > 0 else np.inf, dtype=np.float32)
    return DenFst(N, int(num_pdfs), 0, src.astype(np.int32), dst.astype(np.int32),
                  (pdf + 1).astype(np.int32), cost, final_cost)


def initial_probs(den):
    """Kaldi DenominatorGraph::SetInitialProbs, vectorised; float64."""
    N = den.num_states
    w = np.exp(-den.cost.astype(np.float64))
    tot = np.exp(-den.final_cost.astype(np.float64))
    tot = tot + np.bincount(den.src, weights=w, minlength=N)
    wn = w / tot[den.src]
    cur = np.zeros(N)
    cur[den.start] = 1.0
    avg = np.zeros(N)
    for _ in range(100):
        avg += cur / 100.0
        nxt = np.bincount(den.dst, weights=cur[den.src] * wn, minlength=N)
        cur = nxt / nxt.sum()
    return avg


def make_supervision(seed, den, num_sequences, frames_per_sequence, num_paths=4, weight=1.0,
                     boundary_final_costs=False):
    """Merged supervision FST over num_sequences * frames_per_sequence frames."""
    rng = np.random.default_rng(seed)
    B, T, K = int(num_sequences), int(frames_per_sequence), int(num_paths)
    N = den.num_states
    order = np.argsort(den.src, kind="stable")
    begin = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(den.src, minlength=N), out=begin[1:])
    outdeg = begin[1:] - begin[:-1]
    pi = initial_probs(den)
    cdf = np.cumsum(pi / pi.sum())

    # K walks per sequence, vectorised over (B, K)
    q = np.searchsorted(cdf, rng.random((B, K))).clip(0, N - 1)          # den state at frame 0
    q0 = q.copy()
    arc = np.empty((B, K, T), dtype=np.int64)                             # den arc taken at frame t
    for t in range(T):
        pick = (rng.random((B, K)) * outdeg[q]).astype(np.int64)
        a = order[begin[q] + np.minimum(pick, outdeg[q] - 1)]
        arc[:, :, t] = a
        q = den.dst[a].astype(np.int64)

    den_cost = den.cost.astype(np.float64)
    # merged FST, built sequence by sequence
    arc_src, arc_dst, arc_lab, arc_cost = [], [], [], []
    n_states = 1        # state 0 = global start
    prev_finals = None  # merged ids of the previous sequence's last-level states
    prev_final_cost = None
    for b in range(B):
        # level-1 arcs: keyed by (q0, den arc)
        a0 = arc[b, :, 0]
        key0 = np.unique(np.stack([q0[b], a0], axis=1), axis=0)
        # states per level l = 1..T keyed by den state
        level_ids = []
        level_states = []
        for l in range(1, T + 1):
            ds = np.unique(den.dst[arc[b, :, l - 1]].astype(np.int64))
            ids = n_states + np.arange(ds.shape[0])
            n_states += ds.shape[0]
            level_states.append(ds)
            level_ids.append(ids)
        # start arcs of this sequence: (dst merged id, label, cost)
        d1 = den.dst[key0[:, 1]].astype(np.int64)
        st_dst = level_ids[0][np.searchsorted(level_states[0], d1)]
        st_lab = den.ilabel[key0[:, 1]].astype(np.int64)
        st_cost = den_cost[key0[:, 1]] - np.log(pi[key0[:, 0]])
        if b == 0:
            for j in range(st_dst.shape[0]):
                arc_src.append(0)
                arc_dst.append(int(st_dst[j]))
                arc_lab.append(int(st_lab[j]))
                arc_cost.append(float(st_cost[j]))
        else:
            # Concat + RmEpsilon: every final state f of the previous sequence gets the start arcs,
            # with cost final(f) (x) arc cost
            for f, fc in zip(prev_finals, prev_final_cost):
                for j in range(st_dst.shape[0]):
                    arc_src.append(int(f))
                    arc_dst.append(int(st_dst[j]))
                    arc_lab.append(int(st_lab[j]))
                    arc_cost.append(float(st_cost[j] + fc))
        # inner arcs, level l -> l+1 keyed by den arc id
        for l in range(1, T):
            au = np.unique(arc[b, :, l])
            s_ids = level_ids[l - 1][np.searchsorted(level_states[l - 1], den.src[au].astype(np.int64))]
            d_ids = level_ids[l][np.searchsorted(level_states[l], den.dst[au].astype(np.int64))]
            o = np.argsort(s_ids, kind="stable")
            for j in o:
                arc_src.append(int(s_ids[j]))
                arc_dst.append(int(d_ids[j]))
                arc_lab.append(int(den.ilabel[au[j]]))
                arc_cost.append(float(den_cost[au[j]]))
        prev_finals = level_ids[T - 1]
        if boundary_final_costs:
            prev_final_cost = rng.random(prev_finals.shape[0]) * 2.0
        else:
            prev_final_cost = np.zeros(prev_finals.shape[0])
    S = n_states
    final_cost = np.full(S, np.inf, dtype=np.float32)
    final_cost[prev_finals] = prev_final_cost.astype(np.float32)
    arc_src = np.asarray(arc_src, dtype=np.int64)
    o = np.argsort(arc_src, kind="stable")
    arc_begin = np.zeros(S + 1, dtype=np.int64)
    np.cumsum(np.bincount(arc_src, minlength=S), out=arc_begin[1:])
    return SupFst(float(weight), B, T, den.num_pdfs, arc_begin.astype(np.int32),
                  np.asarray(arc_dst, dtype=np.int32)[o], np.asarray(arc_lab, dtype=np.int32)[o],
                  np.asarray(arc_cost, dtype=np.float32)[o], final_cost)


def append_supervisions(sups):
    """Kaldi AppendSupervision/MergeChainExamples on per-sequence (or multi-sequence)
    supervisions: fst::Concat + RmEpsilon.  The final states of part k get copies of the
    start arcs of part k+1 with cost final(f) (x) arc cost; part k+1's start state vanishes."""
    assert len(sups) >= 1
    T, P, w = sups[0].frames_per_sequence, sups[0].label_dim, sups[0].weight
    src, dst, lab, cost = [], [], [], []
    offset = 0           # merged id of the current part's state 0 (only part 0 keeps it)
    prev_finals = None   # [(merged id, final cost)]
    n_states = 0
    for k, sp in enumerate(sups):
        assert sp.frames_per_sequence == T and sp.label_dim == P
        S = sp.num_states
        # part k's state i>0 maps to base + i; its state 0 maps to merged 0 (k == 0) or disappears
        base = n_states if k == 0 else n_states - 1
        for st in range(S):
            for a in range(sp.arc_begin[st], sp.arc_begin[st + 1]):
                d = base + int(sp.arc_dst[a])
                if st == 0 and k > 0:
                    for f, fc in prev_finals:
                        src.append(f); dst.append(d); lab.append(int(sp.arc_label[a]))
                        cost.append(float(sp.arc_cost[a]) + fc)
                else:
                    src.append(base + st); dst.append(d); lab.append(int(sp.arc_label[a]))
                    cost.append(float(sp.arc_cost[a]))
        prev_finals = [(base + st, float(sp.final_cost[st])) for st in range(S) if np.isfinite(sp.final_cost[st])]
        n_states = base + S
    final_cost = np.full(n_states, np.inf, dtype=np.float32)
    for f, fc in prev_finals:
        final_cost[f] = fc
    src = np.asarray(src, dtype=np.int64)
    o = np.argsort(src, kind="stable")
    arc_begin = np.zeros(n_states + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n_states), out=arc_begin[1:])
    B = sum(sp.num_sequences for sp in sups)
    return SupFst(float(w), B, T, P, arc_begin.astype(np.int32), np.asarray(dst, dtype=np.int32)[o],
                  np.asarray(lab, dtype=np.int32)[o], np.asarray(cost, dtype=np.float32)[o], final_cost)


def interleave_rows(ys, Bs):
    """Frame-major nnet outputs of consecutive sequence groups -> one frame-major matrix."""
    T = ys[0].shape[0] // Bs[0]
    parts = [y.reshape(T, b, -1) for y, b in zip(ys, Bs)]
    return np.concatenate(parts, axis=1).reshape(T * sum(Bs), -1)


def make_nnet_output(seed, num_sequences, frames_per_sequence, num_pdfs, scale=1.0, zero=False):
    """(T*B, P) float32, rows frame-major (row = t*B + b)."""
    rows = int(num_sequences) * int(frames_per_sequence)
    if zero:
        return np.zeros((rows, int(num_pdfs)), dtype=np.float32)
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((rows, int(num_pdfs))) * scale).astype(np.float32)


def write_openfst_binary(path, den):
    """den.fst as OpenFst writes a VectorFst<StdArc> (what ``fst::ReadFstKaldi`` reads at
    /root/reference/src/my_lib_example.cpp:131; layout in SURVEY.md section 8f-1), vectorised: int32 magic,
    "vector" / "standard", version 2, flags 0, properties, start, #states, #arcs, then per state
    {float final, int64 narcs, narcs x {int32 ilabel, int32 olabel, float weight, int32 nextstate}}."""
    import struct
    N, A = den.num_states, den.num_arcs
    order = np.argsort(den.src, kind="stable")
    counts = np.bincount(den.src, minlength=N).astype(np.int64)
    begin = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=begin[1:])
    body = np.zeros(3 * N + 4 * A, dtype=np.uint32)          # everything is a multiple of 4 bytes
    s_off = 3 * np.arange(N, dtype=np.int64) + 4 * begin[:-1]
    body[s_off] = den.final_cost.astype(np.float32).view(np.uint32)
    body[s_off + 1] = (counts & 0xFFFFFFFF).astype(np.uint32)   # int64 narcs, little endian (high word stays 0)
    a_off = np.repeat(s_off + 3, counts) + 4 * (np.arange(A, dtype=np.int64) - np.repeat(begin[:-1], counts))
    body[a_off] = den.ilabel[order].astype(np.int32).view(np.uint32)
    body[a_off + 1] = den.ilabel[order].astype(np.int32).view(np.uint32)
    body[a_off + 2] = den.cost[order].astype(np.float32).view(np.uint32)
    body[a_off + 3] = den.dst[order].astype(np.int32).view(np.uint32)
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 2125659606))
        for t in (b"vector", b"standard"):
            f.write(struct.pack("<i", len(t)) + t)
        f.write(struct.pack("<iiQqqq", 2, 0, 0, den.start, N, A))
        f.write(body.tobytes())


# --- BASELINE.json configurations -------------------------------------------
def small_case(seed):
    """cfg 1: mirrors the ranges of the reference's native test
    (/root/reference/src/chain-supervision-test.hpp:390-399, my_lib_chain.cpp:198-199)."""
    rng = np.random.default_rng(10_000 + seed)
    P = int(rng.integers(5, 61))
    N = int(rng.integers(5, 201))
    B = int(rng.integers(1, 6))
    T = int(rng.integers(2, 21))
    if rng.integers(0, 4) == 0:
        T *= 30
    zero = bool(rng.integers(0, 4) == 0)
    leaky = float([1e-5, 0.2, 0.1][int(rng.integers(0, 3))])
    weight = float([1.0, 0.5][int(rng.integers(0, 2))])
    l2 = float([0.0, 0.01][int(rng.integers(0, 2))])
    K = int(rng.integers(1, 5))
    den = make_den_fst(seed, N, P, mean_deg=float(rng.uniform(0.0, 4.0)), max_deg=4)
    sup = make_supervision(seed + 1, den, B, T, num_paths=K, weight=weight,
                           boundary_final_costs=bool(rng.integers(0, 2)))
    y = make_nnet_output(seed + 2, B, T, P, zero=zero)
    return dict(den=den, sup=sup, y=y, leaky=leaky, l2=l2, B=B, T=T, P=P, N=N)


def add_hubs(den, hubs, seed=0):
    """Heavy-tailed variant of a den graph: `hubs` states get an arc from and to every state
    (real chain den graphs have such states; they exercise the state-splitting path)."""
    rng = np.random.default_rng(seed)
    N, P = den.num_states, den.num_pdfs
    src, dst, lab, cost = [den.src], [den.dst], [den.ilabel], [den.cost]
    for h in range(hubs):
        others = np.arange(N, dtype=np.int32)
        src += [others, np.full(N, h, dtype=np.int32)]
        dst += [np.full(N, h, dtype=np.int32), others]
        lab += [rng.integers(1, P + 1, size=N).astype(np.int32), rng.integers(1, P + 1, size=N).astype(np.int32)]
        cost += [np.full(N, 3.0, dtype=np.float32), np.full(N, np.log(4.0 * N), dtype=np.float32)]
    src, dst, lab, cost = map(np.concatenate, (src, dst, lab, cost))
    o = np.argsort(src, kind="stable")
    return DenFst(N, P, den.start, src[o], dst[o], lab[o], cost[o], den.final_cost)


CONFIGS = {
    # name: (P, N, mean_deg, B, T, leaky, l2)
    "tiny": dict(P=40, N=120, mean_deg=4.0, B=6, T=8, leaky=0.1, l2=5e-5),
    "chime5_hubs": dict(P=2928, N=10_000, mean_deg=26.0, B=128, T=50, leaky=0.1, l2=5e-5, hubs=2),
    "chime5": dict(P=2928, N=10_000, mean_deg=28.0, B=128, T=50, leaky=0.1, l2=5e-5),
    # config 2's shape with pdfs tied to the state being entered, as in a real den.fst (~2 runs per state)
    "chime5_tied": dict(P=2928, N=10_000, mean_deg=28.0, B=128, T=50, leaky=0.1, l2=5e-5, tied=True),
    "chime5_b512": dict(P=2928, N=10_000, mean_deg=28.0, B=512, T=50, leaky=0.1, l2=5e-5),
    "large_tree": dict(P=6000, N=20_000, mean_deg=28.0, B=256, T=50, leaky=0.1, l2=5e-5),
}


def make_config(name, seed=1234, B=None, T=None, y_scale=2.0, with_y=True):
    c = dict(CONFIGS[name])
    if B is not None:
        c["B"] = int(B)
    if T is not None:
        c["T"] = int(T)
    den = (make_den_fst_tied if c.get("tied") else make_den_fst)(seed, c["N"], c["P"], mean_deg=c["mean_deg"])
    sup = make_supervision(seed + 1, den, c["B"], c["T"], num_paths=4, weight=1.0)
    y = make_nnet_output(seed + 2, c["B"], c["T"], c["P"], scale=y_scale) if with_y else None
    c.update(den=den, sup=sup, y=y)
    return c


# ---------------------------------------------------------------------------
# Tied-pdf graphs: in a real den.fst the pdf of an arc belongs to the HMM state
# being entered, so the in-arcs of a state carry very few distinct pdfs.  These helpers build such graphs
# and turn ANY graph into a strictly tied one (exactly one incoming pdf per state) without changing the loss.
# ---------------------------------------------------------------------------
def make_den_fst_tied(seed, num_states, num_pdfs, mean_deg=28.0, final_prob=0.05, untied_frac=0.0):
    """Like make_den_fst, but every arc into state j carries pdf_of[j] (self-loops carry their own
    pdf_loop[j], as in the 2-pdf chain topology); `untied_frac` of the arcs get a random pdf instead."""
    den = make_den_fst(seed, num_states, num_pdfs, mean_deg=mean_deg, final_prob=final_prob)
    rng = np.random.default_rng(seed + 7919)
    N = den.num_states
    pdf_of = rng.integers(0, num_pdfs, size=N)
    pdf_loop = rng.integers(0, num_pdfs, size=N)
    dst = den.dst.astype(np.int64)
    pdf = np.where(den.src == den.dst, pdf_loop[dst], pdf_of[dst])
    if untied_frac > 0:
        odd = rng.random(pdf.shape[0]) < untied_frac
        pdf = np.where(odd, rng.integers(0, num_pdfs, size=pdf.shape[0]), pdf)
    den.ilabel = (pdf + 1).astype(np.int32)
    return den


def in_pdf_runs(den):
    """(number of distinct (state, incoming pdf) pairs, number of states with in-arcs)."""
    key = den.dst.astype(np.int64) * (den.num_pdfs + 1) + den.ilabel.astype(np.int64)
    return int(np.unique(key).shape[0]), int(np.unique(den.dst).shape[0])


def split_states_by_in_pdf(den):
    """Strictly tied equivalent of `den`: state j becomes one copy per distinct incoming pdf q; copy (j, q)
    receives the in-arcs of j labelled q and repeats ALL out-arcs (and the final cost) of j.  Copies of a state
    have identical futures and the copies' forward mass adds up to the original state's, so objf and the
    derivative w.r.t. nnet_output are unchanged (Kaldi's initial-probability iteration included: it is linear).
    Returns (split DenFst, copy_of: original state of every new state)."""
    N, A = den.num_states, den.src.shape[0]
    src, dst, lab = den.src.astype(np.int64), den.dst.astype(np.int64), den.ilabel.astype(np.int64)
    key = dst * (den.num_pdfs + 1) + lab
    uniq, inv = np.unique(key, return_inverse=True)          # one new state per (dst, pdf) pair
    copy_state = uniq // (den.num_pdfs + 1)
    have = np.zeros(N, dtype=bool)
    have[copy_state] = True
    orphans = np.nonzero(~have)[0]                           # states without in-arcs keep one copy
    copy_of = np.concatenate([copy_state, orphans])
    M = copy_of.shape[0]
    # out-arcs of every copy = out-arcs of its original, each pointed at the copy (dst, pdf) it enters
    order = np.argsort(src, kind="stable")
    begin = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=N), out=begin[1:])
    deg = (begin[1:] - begin[:-1])[copy_of]
    new_src = np.repeat(np.arange(M), deg)
    offs = np.concatenate([[0], np.cumsum(deg)])[:-1]
    within = np.arange(int(deg.sum())) - np.repeat(offs, deg)
    arc = order[np.repeat(begin[copy_of], deg) + within]
    new_dst = inv[arc]
    first_copy = np.full(N, -1, dtype=np.int64)
    first_copy[copy_of[::-1]] = np.arange(M)[::-1]           # lowest-numbered copy of every original state
    out = DenFst(int(M), den.num_pdfs, int(first_copy[den.start]), new_src.astype(np.int32), new_dst.astype(np.int32),
                 den.ilabel[arc].astype(np.int32), den.cost[arc].astype(np.float32),
                 den.final_cost[copy_of].astype(np.float32))
    return out, copy_of
