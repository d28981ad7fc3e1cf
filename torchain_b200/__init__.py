"""torchain_b200: B200-native LF-MMI (chain) loss behind torchain's Python API.

``from torchain_b200 import functions, io`` mirrors ``from torchain import
functions, io`` for the chain-loss path.
"""
from . import io  # noqa: F401
from . import functions  # noqa: F401
from .functions import ChainResults, chain_loss, to2d  # noqa: F401

__all__ = ["io", "functions", "ChainResults", "chain_loss", "to2d"]
