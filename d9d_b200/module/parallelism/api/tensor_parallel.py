import torch
import torch.distributed as dist
from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.tensor.parallel import parallelize_module

from d9d_b200.module.block.attention import GroupedQueryAttention
from d9d_b200.module.block.ffn import SwiGLU
from d9d_b200.module.parallelism.style import ColwiseLinearParallel, RowwiseLinearParallel


def parallelize_colwise(module: nn.Linear, mesh: DeviceMesh, tp_dim: str = "tp", sequence_parallel: bool = False) -> None:
    """Column-parallel linear (output features sharded over ``tp_dim``).  Net-new vs the reference."""
    parallelize_module(module, mesh, ColwiseLinearParallel(tp_dim, sequence_parallel))


def parallelize_rowwise(module: nn.Linear, mesh: DeviceMesh, tp_dim: str = "tp", sequence_parallel: bool = False) -> None:
    """Row-parallel linear (input features sharded over ``tp_dim``; outputs all-reduced / reduce-scattered)."""
    parallelize_module(module, mesh, RowwiseLinearParallel(tp_dim, sequence_parallel))


class _GatherAlongSequence:
    """Forward pre-hook: all-gather the rotary tables along the sequence dim so that they cover the tokens of the whole
    tensor-parallel group (inside attention the sequence is complete, outside it is sharded)."""

    def __init__(self, group: dist.ProcessGroup):
        self._group = group

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        parts = [torch.empty_like(t) for _ in range(self._group.size())]
        dist.all_gather(parts, t.contiguous(), group=self._group)
        return torch.cat(parts, dim=1)

    def __call__(self, module: nn.Module, args: tuple, kwargs: dict) -> tuple[tuple, dict]:
        if "position_embeddings" not in kwargs:
            raise ValueError("tensor-parallel attention expects position_embeddings to be passed by keyword")
        cos, sin = kwargs["position_embeddings"]
        with torch.no_grad():
            kwargs = {**kwargs, "position_embeddings": (self._gather(cos), self._gather(sin))}
        return args, kwargs


def parallelize_attention_tensor_parallel(attention: GroupedQueryAttention, mesh: DeviceMesh, tp_dim: str = "tp") -> None:
    """Megatron-style tensor + sequence parallel attention: q / k / v (and the output gate) are column-parallel - every
    rank computes ``heads / tp`` heads over the full sequence - and the output projection is row-parallel.  The block
    takes and returns *sequence-sharded* hidden states ``[B, S / tp, H]`` (contiguous chunks in rank order); on CUDA the
    all-gather / reduce-scatter run inside the GEMMs (``d9d_b200.kernel.tp``).  Per-head q/k norm weights stay replicated
    (their gradients are partial sums over the local heads and are reduced like any replicated parameter)."""
    tp = mesh[tp_dim].size()
    q_out, kv_out = attention.q_proj.weight.shape[0], attention.k_proj.weight.shape[0]
    head_dim = attention.head_dim
    if (q_out // head_dim) % tp != 0 or (kv_out // head_dim) % tp != 0:
        raise ValueError(f"attention heads ({q_out // head_dim} query / {kv_out // head_dim} key-value) must be divisible by "
                         f"the tensor-parallel degree {tp}")
    for name in ("q_proj", "k_proj", "v_proj", "gate_proj"):
        proj = getattr(attention, name)
        if proj is not None:
            parallelize_colwise(proj, mesh, tp_dim, sequence_parallel=True)
    parallelize_rowwise(attention.o_proj, mesh, tp_dim, sequence_parallel=True)
    attention.register_forward_pre_hook(_GatherAlongSequence(mesh.get_group(tp_dim)), with_kwargs=True)


def parallelize_swiglu_tensor_parallel(mlp: SwiGLU, mesh: DeviceMesh, tp_dim: str = "tp") -> None:
    """Column-parallel gate / up projections, row-parallel down projection, sequence-sharded input and output."""
    parallelize_colwise(mlp.gate_proj, mesh, tp_dim, sequence_parallel=True)
    parallelize_colwise(mlp.up_proj, mesh, tp_dim, sequence_parallel=True)
    parallelize_rowwise(mlp.down_proj, mesh, tp_dim, sequence_parallel=True)
