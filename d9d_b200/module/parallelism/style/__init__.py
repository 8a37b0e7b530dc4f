from .shard_experts import ShardMoESparseExpertsParallel
from .tensor_parallel import ColwiseLinearParallel, RowwiseLinearParallel
from .to_local import ToLocalParallel

__all__ = ["ColwiseLinearParallel", "RowwiseLinearParallel", "ShardMoESparseExpertsParallel", "ToLocalParallel"]
