"""User-replaceable data, as in the reference layout (reference ``dataset.py:6-17``: ``FooDataset``; synthetic, seeded
identically on every rank)."""
from b200ddp.data import FooDataset, SyntheticImageNet, SyntheticTokens  # noqa: F401
