from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.tensor.parallel import parallelize_module

from d9d_b200.module.parallelism.style import ColwiseLinearParallel, RowwiseLinearParallel


def parallelize_colwise(module: nn.Linear, mesh: DeviceMesh, tp_dim: str = "tp", sequence_parallel: bool = False) -> None:
    """Column-parallel linear (output features sharded over ``tp_dim``).  Net-new vs the reference."""
    parallelize_module(module, mesh, ColwiseLinearParallel(tp_dim, sequence_parallel))


def parallelize_rowwise(module: nn.Linear, mesh: DeviceMesh, tp_dim: str = "tp", sequence_parallel: bool = False) -> None:
    """Row-parallel linear (input features sharded over ``tp_dim``; outputs all-reduced / reduce-scattered)."""
    parallelize_module(module, mesh, RowwiseLinearParallel(tp_dim, sequence_parallel))
