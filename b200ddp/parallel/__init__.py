from .buckets import BucketSpec, assign_by_size, plan_buckets, bucket_sizes_mib
from .sampler import EndlessSampler, ShardedSampler
from .backend import TorchCollectives, pick_backend_name
from .ddp import DistributedDataParallel
from .data_parallel import DataParallel

DistributedSampler = ShardedSampler

__all__ = ["BucketSpec", "assign_by_size", "plan_buckets", "bucket_sizes_mib", "ShardedSampler", "EndlessSampler", "DistributedSampler",
           "TorchCollectives", "pick_backend_name", "DistributedDataParallel", "DataParallel"]
