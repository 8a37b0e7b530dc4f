from __future__ import annotations

from pathlib import Path

from d9d_b200.core.dist_context import REGULAR_DOMAIN, DistributedContext
from d9d_b200.loop.control import ModelProvider, PrepareExportModelStageContext
from d9d_b200.model_state.io import save_model_state, save_model_state_pipeline_parallel
from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel

from .model_stage_factory import TrackedModules


class ModelStageExporter:
    """Exports the local stages as sharded safetensors through the provider's export mappers."""

    def __init__(self, model_provider: ModelProvider, modules: TrackedModules, dist_context: DistributedContext):
        self._provider, self._modules, self._ctx = model_provider, modules, dist_context

    def export(self, save_dir: Path) -> None:
        mappers = [self._provider.prepare_export_model_stage(PrepareExportModelStageContext(model=m, dist_context=self._ctx)).state_mapper
                   for m in self._modules.modules]
        mapper = ModelStateMapperParallel(mappers)
        if not self._ctx.mesh_params.is_distributed:
            assert len(self._modules.modules) == 1
            save_model_state(dest_dir=save_dir, mapper=mapper, model=self._modules.modules[0], show_progress=self._ctx.is_main_process)
            return
        save_model_state_pipeline_parallel(dest_dir=save_dir, mapper=mapper, device_mesh=self._ctx.mesh_for(REGULAR_DOMAIN),
                                           pipeline_dim_name="pp", models=self._modules.modules, show_progress=True,
                                           position=self._ctx.local_rank)
