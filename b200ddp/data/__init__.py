from .datasets import FooDataset, SyntheticImageNet, SyntheticTokens
from .loader import BatchLoader, DevicePrefetcher

__all__ = ["FooDataset", "SyntheticImageNet", "SyntheticTokens", "BatchLoader", "DevicePrefetcher"]
