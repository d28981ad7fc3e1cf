"""``torchain`` -- the reference's import name, answered by the B200-native implementation.

The reference's callers write ``from torchain import io`` and ``from torchain.functions import
chain_loss, ChainResults`` (/root/reference/test/test.py:6-8,
/root/reference/example/chime5/train_faster.py); this package makes those lines work unmodified on
top of ``torchain_b200`` (same objects, re-exported under the reference's module names).  Like the
reference's own ``torchain/__init__.py`` it imports nothing eagerly.
"""
