"""Independent float64 restatement of the chain objective with dense matrices
and torch autograd -- used only to cross-check oracle/ (which is "parity
unpinned": the reference ships no golden vectors for this path).

Denominator as a linear recursion:
    alpha_0 = pi;   alpha'_t = (I + leaky * pi 1^T) alpha_t
    alpha_t = M_{t-1} alpha'_{t-1},   M_t[h, i] = sum_{arcs i->h} p * exp(y_t[pdf])
    log p_den = log 1^T alpha'_T           (no per-frame normaliser needed in float64 for small T)
Numerator (A.3): log-semiring forward over the merged supervision FST.
Derivatives come from autograd, not from a hand-written backward pass.
"""
import numpy as np
import torch


def den_logprob(den, pi, y, B, T, leaky):
    """y: (T*B, P) float64 torch tensor (requires_grad ok). Returns scalar log-prob summed over sequences."""
    N = den.num_states
    src = torch.as_tensor(den.src.astype(np.int64))
    dst = torch.as_tensor(den.dst.astype(np.int64))
    pdf = torch.as_tensor(den.ilabel.astype(np.int64) - 1)
    p = torch.exp(-torch.as_tensor(den.cost.astype(np.float32))).double()  # float exp as Kaldi
    pi = torch.as_tensor(np.asarray(pi, dtype=np.float64))
    total = y.new_zeros(())
    for s in range(B):
        alpha = pi.clone()
        logscale = y.new_zeros(())
        for t in range(T):
            alpha_dash = alpha + leaky * pi * alpha.sum()
            e = torch.exp(y[t * B + s])
            contrib = alpha_dash[src] * p * e[pdf]
            alpha = torch.zeros(N, dtype=torch.float64).index_add(0, dst, contrib)
            sc = alpha.sum()
            alpha = alpha / sc
            logscale = logscale + torch.log(sc)
        alpha_dash = alpha + leaky * pi * alpha.sum()
        total = total + torch.log(alpha_dash.sum()) + logscale
    return total


def num_logprob(sup, y):
    """Log-semiring forward over the merged FST; returns unweighted total log-prob."""
    B, T = sup.num_sequences, sup.frames_per_sequence
    S = sup.num_states
    ninf = float("-inf")
    la = [None] * S
    la[0] = y.new_zeros(())
    time = [-1] * S
    time[0] = 0
    tot = None
    for s in range(S):
        if la[s] is None:
            continue
        t = time[s]
        for a in range(sup.arc_begin[s], sup.arc_begin[s + 1]):
            d = int(sup.arc_dst[a])
            seq, frame = t // T, t % T
            v = la[s] - float(sup.arc_cost[a]) + y[frame * B + seq, int(sup.arc_label[a]) - 1]
            la[d] = v if la[d] is None else torch.logaddexp(la[d], v)
            time[d] = t + 1
        if np.isfinite(sup.final_cost[s]):
            v = la[s] - float(sup.final_cost[s])
            tot = v if tot is None else torch.logaddexp(tot, v)
    return tot


def chain_objf(den, pi, sup, y_np, leaky, l2):
    """Returns (objf, l2_term, weight, deriv) in float64 with autograd derivatives of objf + l2_term."""
    y = torch.tensor(np.asarray(y_np, dtype=np.float64), requires_grad=True)
    B, T = sup.num_sequences, sup.frames_per_sequence
    w = float(sup.weight)
    num = num_logprob(sup, y)
    denlp = den_logprob(den, pi, y, B, T, leaky)
    objf = w * num - w * denlp
    l2_term = -0.5 * w * l2 * (y * y).sum()
    (objf + l2_term).backward()
    return float(objf.detach()), float(l2_term.detach()), w * B * T, y.grad.numpy().copy()


def den_logprob_batched(den, pi, y, B, T, leaky):
    """The same linear recursion as den_logprob with all B sequences advanced together, so that
    it can run on graphs of BASELINE.json's real sizes (N = 10 000 states, A ~ 300 000 arcs): the
    per-arc products are one (A, B) array per frame and the scatter is torch's index_add.  It
    shares no code with oracle/ (no arc loops, no hand-written backward pass, no Kaldi-style
    normaliser bookkeeping: derivatives are autograd's), which is what makes it a second,
    independent formulation at realistic size.  y: (T*B, P) float64, rows frame-major."""
    N = den.num_states
    src = torch.as_tensor(den.src.astype(np.int64))
    dst = torch.as_tensor(den.dst.astype(np.int64))
    pdf = torch.as_tensor(den.ilabel.astype(np.int64) - 1)
    p = torch.exp(-torch.as_tensor(den.cost.astype(np.float32))).double()[:, None]  # float exp as Kaldi
    pi = torch.as_tensor(np.asarray(pi, dtype=np.float64))[:, None]
    e = torch.exp(y).reshape(T, B, -1)
    alpha = pi.expand(N, B)
    logscale = y.new_zeros(B)
    for t in range(T):
        alpha_dash = alpha + leaky * pi * alpha.sum(0, keepdim=True)
        contrib = alpha_dash[src] * p * e[t].t()[pdf]
        alpha = torch.zeros(N, B, dtype=torch.float64).index_add(0, dst, contrib)
        sc = alpha.sum(0)
        alpha = alpha / sc
        logscale = logscale + torch.log(sc)
    alpha_dash = alpha + leaky * pi * alpha.sum(0, keepdim=True)
    return (torch.log(alpha_dash.sum(0)) + logscale).sum()


def chain_objf_batched(den, pi, sup, y_np, leaky, l2):
    """chain_objf with the batched denominator: usable on the full-size graphs at small B."""
    y = torch.tensor(np.asarray(y_np, dtype=np.float64), requires_grad=True)
    B, T = sup.num_sequences, sup.frames_per_sequence
    w = float(sup.weight)
    objf = w * num_logprob(sup, y) - w * den_logprob_batched(den, pi, y, B, T, leaky)
    l2_term = -0.5 * w * l2 * (y * y).sum()
    (objf + l2_term).backward()
    return float(objf.detach()), float(l2_term.detach()), w * B * T, y.grad.numpy().copy()
