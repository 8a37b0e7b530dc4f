"""Distributed model-state IO and context helpers on gloo (reference ``test/d9d_test/model_state/test_dist_io.py``,
``core/test_dist_context.py``): sharded multi-writer save, pipeline-parallel save with one writer per stage,
exact ``model.safetensors.index.json`` contents, round trips, ``main_process_first`` ordering, timeouts, batch maths."""

import json
import time
from pathlib import Path

import pytest
import torch
from torch import nn

from tests.dist_utils import run_distributed


def _read_back(path: Path, names: list[str]) -> dict[str, torch.Tensor]:
    from d9d_b200.model_state.io import read_model_state
    from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel
    from d9d_b200.model_state.mapper.leaf import ModelStateMapperIdentity

    mapper = ModelStateMapperParallel([ModelStateMapperIdentity(n) for n in names])
    return dict(read_model_state(path, mapper, device="cpu", show_progress=False))


# ---------------------------------------------------------------------------------- every rank writes its share
def _sharded_writer(rank, world_size, tmp):
    import torch.distributed as dist

    from d9d_b200.model_state.io import write_model_state_distributed
    from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel, ModelStateMapperShard
    from d9d_b200.model_state.mapper.leaf import ModelStateMapperIdentity

    dist.init_process_group("gloo")
    names = [f"layer.{i}.weight" for i in range(6)]
    state = {n: torch.full((4, 4), float(i)) for i, n in enumerate(names)}  # 64 B each
    everything = ModelStateMapperParallel([ModelStateMapperIdentity(n) for n in names])
    mine = ModelStateMapperShard(everything, total_shards=world_size, current_shard=rank)
    write_model_state_distributed(Path(tmp), mine, iter(state.items()), process_group=dist.group.WORLD, show_progress=False)
    dist.barrier()
    if rank == 0:
        index = json.loads((Path(tmp) / "model.safetensors.index.json").read_text())
        assert index["metadata"]["total_size"] == 6 * 64
        assert sorted(index["weight_map"]) == names
        files = sorted(set(index["weight_map"].values()))
        assert files == [f"model-{i + 1:05d}-of-{len(files):05d}.safetensors" for i in range(len(files))]
        assert sorted(p.name for p in Path(tmp).iterdir()) == sorted([*files, "model.safetensors.index.json"])  # no temp files left
        back = _read_back(Path(tmp), names)
        for n in names:
            assert torch.equal(back[n], state[n])


@pytest.mark.dist
def test_sharded_distributed_save(tmp_path):
    run_distributed(_sharded_writer, 2, str(tmp_path))


# ------------------------------------------------------------------------------ one writer per pipeline stage
class _Stage(nn.Module):
    def __init__(self, index: int):
        super().__init__()
        self.add_module(f"block_{index}", nn.Linear(3, 3, bias=False))
        with torch.no_grad():
            self.get_submodule(f"block_{index}").weight.fill_(float(index))


def _pp_writer(rank, world_size, tmp):
    from d9d_b200.core.dist_context import REGULAR_DOMAIN, DeviceMeshParameters
    from d9d_b200.model_state.io import save_model_state_pipeline_parallel
    from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel
    from d9d_b200.model_state.mapper.leaf import ModelStateMapperIdentity

    ctx = DeviceMeshParameters(pipeline_parallel=2, data_parallel_replicate=2).build()
    mesh = ctx.mesh_for(REGULAR_DOMAIN)
    pp = mesh.get_local_rank("pp")
    names = ["block_0.weight", "block_1.weight"]
    mapper = ModelStateMapperParallel([ModelStateMapperIdentity(names[pp])])  # each stage exports the states it holds
    save_model_state_pipeline_parallel(Path(tmp), mapper, device_mesh=mesh, pipeline_dim_name="pp", models=[_Stage(pp)],
                                       show_progress=False)
    ctx.wait_world()
    if ctx.is_main_process:
        index = json.loads((Path(tmp) / "model.safetensors.index.json").read_text())
        assert sorted(index["weight_map"]) == names  # replicas of a stage did not write duplicates
        assert index["metadata"]["total_size"] == 2 * 9 * 4
        assert len(list(Path(tmp).glob("*.safetensors"))) == 2
        back = _read_back(Path(tmp), names)
        assert back["block_0.weight"].eq(0).all() and back["block_1.weight"].eq(1).all()


@pytest.mark.dist
def test_pipeline_parallel_save_has_one_writer_per_stage(tmp_path):
    run_distributed(_pp_writer, 4, str(tmp_path))


# ------------------------------------------------------------------------------------------- context helpers
def _context_helpers(rank, world_size, tmp):
    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.component.batch_maths import BatchMaths
    from d9d_b200.loop.component.timeout_manager import TimeoutManager
    from d9d_b200.loop.config import BatchingConfig
    from d9d_b200.loop.config.config import TimeoutConfig

    ctx = DeviceMeshParameters(data_parallel_replicate=2).build()
    marker = Path(tmp) / "prepared"
    with ctx.main_process_first():
        if ctx.is_main_process:
            time.sleep(0.3)
            marker.write_text("done")
        else:
            assert marker.exists()  # the main process finished its block before anyone else entered
    with ctx.local_main_process_first():
        assert marker.exists()

    timeouts = TimeoutManager(ctx, TimeoutConfig(init_timeout=600, step_timeout=120))
    with pytest.raises(ValueError):
        timeouts.set_periodic()  # the init timeout has to be armed first
    timeouts.set_init()
    timeouts.set_periodic()
    timeouts.set_periodic()  # idempotent once steps are flowing
    with pytest.raises(ValueError):
        timeouts.set_init()

    maths = BatchMaths(ctx, BatchingConfig(global_batch_size=16, microbatch_size=2), None)
    assert maths.data_parallel_size == 2
    assert maths.num_microbatches_gradient_accumulation == 4 and maths.num_microbatches_pipelining == 1
    assert maths.data_loader_batch_size == 2 and maths.num_backward_calls == 4
    with pytest.raises(ValueError):
        BatchMaths(ctx, BatchingConfig(global_batch_size=6, microbatch_size=2), None)

    assert ctx.logger.name.startswith("d9d") or "d9d" in ctx.logger.name
    assert ctx.num_nodes == 1 and ctx.node_rank == 0


@pytest.mark.dist
def test_context_helpers_timeouts_and_batch_maths(tmp_path):
    run_distributed(_context_helpers, 2, str(tmp_path))


def _pp_batch_maths(rank, world_size):
    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.component.batch_maths import BatchMaths
    from d9d_b200.loop.config import BatchingConfig

    ctx = DeviceMeshParameters(pipeline_parallel=2).build()
    maths = BatchMaths(ctx, BatchingConfig(global_batch_size=8, microbatch_size=2), None)
    # with pipelining the loader yields the whole step's batch and the schedule splits it into microbatches
    assert maths.num_microbatches_pipelining == 4 and maths.num_microbatches_gradient_accumulation == 1
    assert maths.data_loader_batch_size == 8 and maths.num_backward_calls == 4


@pytest.mark.dist
def test_batch_maths_with_pipeline_parallelism():
    run_distributed(_pp_batch_maths, 2)
