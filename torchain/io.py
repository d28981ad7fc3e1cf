"""``torchain.io`` (/root/reference/torchain/io.py): set_kaldi_device, Supervision, DenominatorGraph
for the chain-loss path.  The Kaldi egs readers of that module (Example, RandExample, open_example,
print_key_length; io.py:59-175) are outside the hot path: the names exist and explain themselves."""
from torchain._ext import my_lib
from torchain_b200.io import DenominatorGraph, Supervision, set_kaldi_device  # noqa: F401

ffi = my_lib._my_lib.ffi


def cstr(s):
    return ffi.new("char[]", s.encode())


def _egs_reader(name):
    def ctor(*args, **kwargs):
        raise NotImplementedError(
            "torchain.io.%s reads Kaldi nnet3 chain egs through Kaldi itself, which is outside the chain-loss "
            "hot path this library replaces; build minibatch supervisions with Supervision.from_arrays, "
            "Supervision.from_kaldi_bytes / from_chain_egs and Supervision.merge" % name)
    ctor.__name__ = name
    return ctor


Example = _egs_reader("Example")
RandExample = _egs_reader("RandExample")
open_example = _egs_reader("open_example")
print_key_length = _egs_reader("print_key_length")
