from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class VocabParallelOptions:
    start: int
    stop: int
    group: dist.ProcessGroup | None = None
    reduce_e_grad: bool = True


def vp_reduce_lse(lse: torch.Tensor, pg) -> torch.Tensor:
    world = dist.get_world_size(pg)
    gathered = [torch.empty_like(lse) for _ in range(world)]
    dist.all_gather(gathered, lse, group=pg)
    return torch.logsumexp(torch.stack(gathered), dim=0)


def vp_reduce_correct_logit(neg_correct_logit: torch.Tensor, pg, dtype: torch.dtype | None = None) -> torch.Tensor:
    out = neg_correct_logit.to(dtype) if dtype is not None else neg_correct_logit.clone()
    dist.all_reduce(out, group=pg)
    return out
