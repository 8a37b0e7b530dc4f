from __future__ import annotations

from collections.abc import Iterator
from pathlib import Path

import torch
from safetensors import safe_open
from tqdm import tqdm

from d9d_b200.model_state.mapper import ModelStateMapper, StateGroup

from .dto import MODEL_STATE_INDEX_FILE_NAME, ModelStateIndex


def read_model_state(src_dir: Path, mapper: ModelStateMapper, device: str, show_progress: bool = True,
                     position: int | None = None) -> Iterator[tuple[str, torch.Tensor]]:
    """Stream a sharded safetensors checkpoint through ``mapper``.

    Shard files are opened one by one; only the tensors some dependency group needs are read; a group fires as soon
    as all its inputs are resident and its inputs are evicted right after, so peak memory ~ one shard + the largest
    group.  Yields ``(output name, tensor)``.
    """
    src_dir = Path(src_dir)
    index = ModelStateIndex.model_validate_json((src_dir / MODEL_STATE_INDEX_FILE_NAME).read_text(encoding="utf-8"))

    groups: list[StateGroup] = list(mapper.state_dependency_groups())
    wanted: set[str] = set()
    for g in groups:
        wanted |= g.inputs
    missing = wanted - index.weight_map.keys()
    if missing:
        raise ValueError(f"Cannot run state loading: states {missing} are missing!")

    # file -> keys to read, in first-seen order; group bookkeeping by outstanding input count
    plan: dict[str, list[str]] = {}
    for key in sorted(wanted):
        plan.setdefault(index.weight_map[key], []).append(key)
    waiting: dict[str, list[int]] = {}
    remaining = []
    for gi, g in enumerate(groups):
        remaining.append(len(g.inputs))
        for key in g.inputs:
            waiting.setdefault(key, []).append(gi)

    resident: dict[str, torch.Tensor] = {}
    desc = f"Loading Model States [{position}]" if position is not None else "Loading Model States"
    with tqdm(desc=desc, total=sum(len(g.outputs) for g in groups), disable=not show_progress, position=position, leave=True) as bar:
        for gi, g in enumerate(groups):  # groups without inputs fire immediately
            if remaining[gi] == 0:
                produced = mapper.apply({})
                bar.update(len(produced))
                yield from produced.items()
        for file_name, keys in plan.items():
            ready: list[int] = []
            with safe_open(str(src_dir / file_name), framework="pt", device=str(device)) as shard:
                for key in keys:
                    resident[key] = shard.get_tensor(key)
                    for gi in waiting.get(key, ()):
                        remaining[gi] -= 1
                        if remaining[gi] == 0:
                            ready.append(gi)
            for gi in ready:
                g = groups[gi]
                produced = mapper.apply({k: resident[k] for k in g.inputs})
                for k in g.inputs:
                    resident.pop(k, None)
                bar.update(len(produced))
                yield from produced.items()
