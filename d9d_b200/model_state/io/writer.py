from __future__ import annotations

import warnings
from collections.abc import Iterable
from pathlib import Path

import torch
from safetensors.torch import save_file
from torch.distributed import DeviceMesh, ProcessGroup
from tqdm import tqdm

from d9d_b200.core.dist_ops import all_gather_object
from d9d_b200.model_state.mapper import ModelStateMapper, StateGroup

from .dto import MODEL_STATE_INDEX_FILE_NAME, ModelStateIndex, ModelStateIndexMeta


class _ShardWriter:
    """Buffers mapped tensors and spills them into ``.tmp-rank{r}-shard-{k}.safetensors`` files of bounded size."""

    def __init__(self, dest_dir: Path, rank_tag: int, shard_size_bytes: int):
        self._dir, self._tag, self._limit = dest_dir, rank_tag, shard_size_bytes
        self._pending: dict[str, torch.Tensor] = {}
        self._pending_bytes = 0
        self._files: list[str] = []
        self.weight_map: dict[str, str] = {}
        self.total_bytes = 0

    def add(self, name: str, tensor: torch.Tensor) -> None:
        size = tensor.numel() * tensor.element_size()
        if size > self._limit:
            raise ValueError(f"Cannot save state {name} that is larger than shard size")
        if self._pending_bytes + size > self._limit:
            self.flush()
        self._pending[name] = tensor
        self._pending_bytes += size

    def flush(self) -> int:
        if not self._pending:
            return 0
        file_name = f".tmp-rank{self._tag}-shard-{len(self._files) + 1}.safetensors"
        save_file({k: v.detach().contiguous().cpu() if v.device.type != "cpu" else v.detach().contiguous() for k, v in self._pending.items()},
                  str(self._dir / file_name))
        self._files.append(file_name)
        for k in self._pending:
            self.weight_map[k] = file_name
        n = len(self._pending)
        self.total_bytes += self._pending_bytes
        self._pending.clear()
        self._pending_bytes = 0
        return n


def _stream_write(dest_dir: Path, mapper: ModelStateMapper, states: Iterable[tuple[str, torch.Tensor]], shard_size_gb: float,
                  show_progress: bool, rank_tag: int, writes: bool, position: int | None) -> ModelStateIndex | None:
    """Feed ``states`` through ``mapper`` group by group.  Every rank runs the mapper (it may contain collectives such
    as ``full_tensor``); only ranks with ``writes`` buffer and persist the results."""
    dest_dir.mkdir(parents=True, exist_ok=True)
    groups: list[StateGroup] = list(mapper.state_dependency_groups())
    remaining = [len(g.inputs) for g in groups]
    waiting: dict[str, list[int]] = {}
    for gi, g in enumerate(groups):
        for key in g.inputs:
            waiting.setdefault(key, []).append(gi)
    writer = _ShardWriter(dest_dir, rank_tag, int(shard_size_gb * (1024**3)))
    available: dict[str, torch.Tensor] = {}
    unused: list[str] = []
    done = [False] * len(groups)
    desc = f"Saving Model States [{position}]" if position is not None else "Saving Model States"
    with tqdm(desc=desc, total=sum(len(g.outputs) for g in groups), disable=not (show_progress and writes), position=position, leave=True) as bar:
        for name, tensor in states:
            if name not in waiting:
                unused.append(name)
                continue
            available[name] = tensor
            for gi in waiting[name]:
                remaining[gi] -= 1
                if remaining[gi] == 0 and not done[gi]:
                    done[gi] = True
                    g = groups[gi]
                    produced = mapper.apply({k: available[k] for k in g.inputs})
                    for k in g.inputs:
                        available.pop(k, None)
                    if writes:
                        for out_name, out_tensor in produced.items():
                            writer.add(out_name, out_tensor)
                        bar.update(len(produced))
        if not writes:
            return None
        writer.flush()
    if not all(done):
        missing = {groups[i].inputs for i, d in enumerate(done) if not d}
        raise ValueError("Writing failed: not all source tensors were provided to satisfy mapper dependencies. "
                         f"Missing inputs for groups: {missing}")
    if unused:
        warnings.warn("State Writing: The following source tensors were provided but not consumed by any mapper group "
                      f"and will be ignored: {sorted(unused)}", stacklevel=2)
    return ModelStateIndex(metadata=ModelStateIndexMeta(total_size=writer.total_bytes), weight_map=writer.weight_map)


def _finalize(dest_dir: Path, indices: list[ModelStateIndex]) -> None:
    """Rename temporary shards to ``model-{i:05d}-of-{n:05d}.safetensors`` and write the global index."""
    merged: dict[str, str] = {}
    for idx in indices:
        merged.update(idx.weight_map)
    tmp_files = list(dict.fromkeys(merged.values()))
    final_of = {tmp: f"model-{i + 1:05d}-of-{len(tmp_files):05d}.safetensors" for i, tmp in enumerate(tmp_files)}
    for tmp, final in final_of.items():
        (dest_dir / tmp).rename(dest_dir / final)
    index = ModelStateIndex(metadata=ModelStateIndexMeta(total_size=sum(i.metadata.total_size for i in indices)),
                            weight_map={name: final_of[tmp] for name, tmp in merged.items()})
    (dest_dir / MODEL_STATE_INDEX_FILE_NAME).write_text(index.model_dump_json(indent=4), encoding="utf-8")


def write_model_state_local(dest_dir: Path, mapper: ModelStateMapper, state_generator: Iterable[tuple[str, torch.Tensor]],
                            shard_size_gb: float = 4.0, show_progress: bool = True) -> None:
    """Single-process save."""
    dest_dir = Path(dest_dir)
    idx = _stream_write(dest_dir, mapper, state_generator, shard_size_gb, show_progress, 0, True, None)
    assert idx is not None
    _finalize(dest_dir, [idx])


def write_model_state_distributed(dest_dir: Path, mapper: ModelStateMapper, state_generator: Iterable[tuple[str, torch.Tensor]],
                                  process_group: ProcessGroup, shard_size_gb: float = 4.0, show_progress: bool = True,
                                  position: int | None = None) -> None:
    """Every rank writes its own shards (e.g. with a ``ModelStateMapperShard``); group rank 0 merges the indices."""
    dest_dir = Path(dest_dir)
    idx = _stream_write(dest_dir, mapper, state_generator, shard_size_gb, show_progress, process_group.rank(), True, position)
    gathered = [i for i in all_gather_object(idx, process_group) if i is not None]
    if process_group.rank() == 0:
        _finalize(dest_dir, gathered)


def write_model_state_pipeline_parallel(dest_dir: Path, mapper: ModelStateMapper, state_generator: Iterable[tuple[str, torch.Tensor]],
                                        device_mesh: DeviceMesh, pipeline_dim_name: str, shard_size_gb: float = 4.0,
                                        show_progress: bool = True, position: int | None = None) -> None:
    """N-D mesh save: per pipeline rank only the process whose non-pp coordinates are all zero writes."""
    dest_dir = Path(dest_dir)
    names, coords = device_mesh.mesh_dim_names, device_mesh.get_coordinate()
    if names is None or coords is None:
        raise ValueError("Cannot save state using a DeviceMesh with no dim names or coords")
    pp_rank = device_mesh[pipeline_dim_name].get_local_rank()
    is_writer = all(c == 0 for n, c in zip(names, coords, strict=True) if n != pipeline_dim_name)
    idx = _stream_write(dest_dir, mapper, state_generator, shard_size_gb, show_progress, pp_rank, is_writer, position)
    world_group = _whole_mesh_group(device_mesh)
    gathered = [i for i in all_gather_object(idx, world_group) if i is not None]
    if pp_rank == 0 and is_writer:
        _finalize(dest_dir, gathered)


def _whole_mesh_group(device_mesh: DeviceMesh) -> ProcessGroup:
    """A group spanning every rank of the mesh (the default group when the mesh covers the world)."""
    import torch.distributed as dist

    if device_mesh.size() == dist.get_world_size():
        return dist.group.WORLD
    if device_mesh.ndim == 1:
        return device_mesh.get_group(0)
    return device_mesh._flatten().get_group()  # noqa: SLF001
