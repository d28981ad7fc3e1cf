"""``torchain.functions`` (/root/reference/torchain/functions.py): ChainResults, _ChainLoss, to2d,
chain_loss -- the B200-native implementations of torchain_b200.functions under the reference's name."""
from torchain import io  # noqa: F401  (the reference imports it here too)
from torchain._ext import my_lib  # noqa: F401
from torchain_b200.functions import ChainResults, _ChainLoss, chain_loss, to2d  # noqa: F401

ffi = my_lib._ffi

__all__ = ["ChainResults", "_ChainLoss", "chain_loss", "to2d"]
