"""The reference's one integration test, on synthetic data.

/root/reference/test/test.py:19-78 trains a 3-layer Conv1d model with an LF-MMI head and a cross-entropy
head for 11 SGD steps with ``chain_loss(pred, den_graph, supervision, l2_regularize=0.01,
xent_regularize=0.01, xent_input=xent, kaldi_way=True)`` on CHiME-5 egs and prints ``ChainResults``; it asserts
nothing.  This module keeps the model and the loop (same kernel 29 / stride 3 front end, same loss arguments,
the model's (B, C, T) outputs handed to the loss as they are) and adds what can be asserted on any data.  The
loss function is a parameter: the GPU suite passes ``torchain_b200.functions.chain_loss``, the CPU suite drives
the same loop with an oracle-backed autograd function to pin the expectations themselves.
"""
import numpy as np
import torch
import torch.nn as nn

from torchain_b200 import synth

N_FEAT, HIDDEN = 40, 48


class Model(nn.Module):
    """Reference: test/test.py:19-32 (512 hidden units there)."""

    def __init__(self, n_pdf):
        super().__init__()
        self.common = nn.Sequential(nn.Conv1d(N_FEAT, HIDDEN, 29, 3), nn.ReLU(), nn.Conv1d(HIDDEN, HIDDEN, 1, 1))
        self.lf_mmi = nn.Conv1d(HIDDEN, n_pdf, 1, 1)
        self.xent = nn.Conv1d(HIDDEN, n_pdf, 1, 1)

    def forward(self, x):
        h = self.common(x)
        return self.lf_mmi(h), self.xent(h)


def make_task(seed=0, n_state=120, n_pdf=30, B=6, T=8):
    den = synth.make_den_fst(seed, n_state, n_pdf, mean_deg=3.0)
    sup = synth.make_supervision(seed + 1, den, B, T, num_paths=2, weight=1.0)
    gen = torch.Generator().manual_seed(seed + 2)
    x = torch.randn(B, N_FEAT, 29 + 3 * (T - 1), generator=gen)  # T output frames after kernel 29 / stride 3
    return den, sup, x


def run_sgd(loss_fn, den_graph, supervision, x, n_pdf, device, steps=11, lr=1e-3, l2=0.01, xent=0.01):
    """The loop of test/test.py:57-75 on one fixed minibatch; returns per step
    (loss, [objf, l2_term, weight], |grad| of the LF-MMI head, |grad| of the xent head)."""
    torch.manual_seed(1234)
    model = Model(n_pdf).to(device)
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    x = x.to(device)
    hist = []
    for _ in range(steps):
        pred, xe = model(x)
        loss, results = loss_fn(pred, den_graph, supervision, l2_regularize=l2, xent_regularize=xent,
                                xent_input=xe, kaldi_way=True)
        opt.zero_grad()
        loss.backward()
        hist.append((float(loss.detach()), results.data.clone().numpy().astype(np.float64),
                     float(model.lf_mmi.weight.grad.abs().sum()), float(model.xent.weight.grad.abs().sum())))
        opt.step()
    return hist


def check_history(hist, B, T):
    for loss, data, g_mmi, g_xent in hist:
        objf, l2_term, weight = data
        assert np.isfinite([loss, objf, l2_term, weight, g_mmi, g_xent]).all()
        assert weight == B * T                       # supervision.weight * B * T
        assert objf <= 0.0 and l2_term <= 0.0        # chain-supervision-test.hpp:285
        assert abs(loss - (-objf / weight)) <= 1e-5 * max(1.0, abs(loss))   # functions.py:17-19, no l2 term
        assert g_mmi > 0.0 and g_xent > 0.0          # both heads receive a gradient (kaldi_way xent output)
    # backward() hands out -(d objf + d l2_term), un-normalised (functions.py:106-115): SGD on a fixed
    # minibatch must have reduced -(objf + l2_term) after 11 steps
    first, last = hist[0][1], hist[-1][1]
    assert -(last[0] + last[1]) < -(first[0] + first[1])
