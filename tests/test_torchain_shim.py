"""The reference's import name: ``from torchain import io``, ``from torchain.functions import chain_loss,
ChainResults``, ``from torchain._ext import my_lib`` (/root/reference/test/test.py:6-8,
/root/reference/torchain/functions.py:1-6) must work unmodified on top of the B200 library."""
import inspect

import numpy as np
import pytest

import oracle_lib as ol
from common import DERIV_RTOL, OBJF_RTOL, deriv_err, rel
from torchain_b200 import synth


def test_reference_import_lines_and_signatures():
    from torchain import io
    from torchain.functions import ChainResults, _ChainLoss, chain_loss, to2d  # noqa: F401
    from torchain._ext import my_lib
    import torchain_b200
    assert chain_loss is torchain_b200.functions.chain_loss and io.Supervision is torchain_b200.io.Supervision
    sig = inspect.signature(chain_loss)
    assert list(sig.parameters) == ["input", "den_graph", "supervision", "l2_regularize", "leaky_hmm_coefficient",
                                    "xent_regularize", "xent_input", "kaldi_way"]      # functions.py:128-130
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d == dict(l2_regularize=0.0, leaky_hmm_coefficient=1e-5, xent_regularize=0.0, xent_input=None,
                     kaldi_way=False)
    r = ChainResults()
    assert tuple(r.data.shape) == (3,) and not r.data.is_cuda
    ffi = my_lib._ffi
    assert ffi is my_lib._my_lib.ffi and not ffi.NULL and ffi.NULL == ffi.NULL
    assert io.cstr("den.fst").value == b"den.fst"
    # reference io.py:23-24: a NULL handle from the native layer is a ValueError
    with pytest.raises(ValueError, match="null supervision ptr"):
        io.Supervision(my_lib.my_lib_supervision_new(ffi.NULL))
    with pytest.raises(NotImplementedError):
        io.RandExample("train.scp", 1, 128)
    with pytest.raises(NotImplementedError):
        my_lib.my_lib_example_reader_new(io.cstr("ark:egs"))


def _handles(io, den, sup):
    g = io.DenominatorGraph.from_arrays(den.num_states, den.src, den.dst, den.ilabel, den.cost, den.start,
                                        den.final_cost, den.num_pdfs)
    s = io.Supervision.from_arrays(sup.weight, sup.num_sequences, sup.frames_per_sequence, sup.label_dim,
                                   sup.arc_begin, sup.arc_dst, sup.arc_label, sup.arc_cost, sup.final_cost)
    return g, s


@pytest.mark.gpu
def test_reference_call_pattern_through_torchain_name():
    """test/test.py:66-74 as written there: (B, C, T) predictions of two heads, chain_loss with xent, backward,
    results accumulated with `+=` (train_faster.py:145)."""
    import torch
    from torchain import io
    from torchain.functions import ChainResults, chain_loss
    c = synth.small_case(14)
    B, T, P = c["B"], c["T"], c["P"]
    den_graph, supervision = _handles(io, c["den"], c["sup"])
    assert supervision.shape == (B, T, P)
    r = ol.chain(ol.OracleDen(c["den"]), ol.OracleSup(c["sup"]), c["y"], l2=0.01, leaky=1e-5, mode="fp64",
                 want_xent=True)
    pred = torch.from_numpy(c["y"]).view(T, B, P).permute(1, 2, 0).contiguous().cuda().requires_grad_(True)
    xent = pred.detach().clone().requires_grad_(True)
    train_result = ChainResults()
    loss, results = chain_loss(pred, den_graph, supervision, l2_regularize=0.01,
                               xent_regularize=0.01, xent_input=xent, kaldi_way=True)
    loss.backward()
    train_result.data += results.data
    assert rel(float(train_result.data[0]), r["objf"]) < OBJF_RTOL
    assert float(train_result.loss) == pytest.approx(-r["objf"] / r["weight"], rel=1e-5)
    gy = pred.grad.permute(2, 0, 1).reshape(T * B, P).cpu().numpy()
    gx = xent.grad.permute(2, 0, 1).reshape(T * B, P).cpu().numpy()
    assert deriv_err(gy, -r["deriv"]) < DERIV_RTOL
    assert deriv_err(gx, -0.01 * r["xent"]) < DERIV_RTOL


@pytest.mark.gpu
def test_my_lib_object_takes_tensors_like_the_cffi_module():
    """functions.py:75-86 verbatim: tensors and ffi.NULL straight into my_lib.my_lib_ComputeChainObjfAndDeriv."""
    import torch
    from torchain import io
    from torchain._ext import my_lib
    ffi = my_lib._ffi
    c = synth.small_case(21)
    den_graph, supervision = _handles(io, c["den"], c["sup"])
    r = ol.chain(ol.OracleDen(c["den"]), ol.OracleSup(c["sup"]), c["y"], l2=c["l2"], leaky=c["leaky"], mode="fp64",
                 want_xent=True)
    my_lib.my_lib_set_kaldi_device(torch.zeros(1).cuda(0))
    input = torch.from_numpy(c["y"]).cuda()
    results = torch.zeros(3)
    mmi_grad = input.new(*input.shape)
    assert my_lib.my_lib_ComputeChainObjfAndDeriv(den_graph.ptr, supervision.ptr, input, results, mmi_grad, ffi.NULL,
                                                  c["l2"], c["leaky"], 0.0) == 1
    assert rel(float(results[0]), r["objf"]) < OBJF_RTOL
    assert deriv_err(mmi_grad.cpu().numpy(), r["deriv"]) < DERIV_RTOL
    xent_grad = input.new()                  # empty, resized by the callee (src/my_lib_chain.cpp:22-29)
    assert my_lib.my_lib_ComputeChainObjfAndDeriv(den_graph.ptr, supervision.ptr, input, results, mmi_grad, xent_grad,
                                                  c["l2"], c["leaky"], 0.1) == 1
    assert deriv_err(xent_grad.cpu().numpy(), r["xent"]) < DERIV_RTOL
    with pytest.raises(RuntimeError, match="dim != 2"):
        my_lib.my_lib_ComputeChainObjfAndDeriv(den_graph.ptr, supervision.ptr, input.view(-1), results, mmi_grad,
                                               ffi.NULL, 0.0, 0.1, 0.0)
