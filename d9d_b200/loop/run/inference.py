from __future__ import annotations

import torch
from tqdm import tqdm

from d9d_b200.core.dist_context import DeviceMeshParameters
from d9d_b200.internals.pipeline_state import PipelineStateHandler
from d9d_b200.loop import component as parts
from d9d_b200.loop import control
from d9d_b200.loop.config import InferenceConfig, PipeliningConfig
from d9d_b200.loop.event.catalogue import common as shared_events
from d9d_b200.loop.event.catalogue import inference as events
from d9d_b200.loop.state import InferenceJobState
from d9d_b200.pipelining.factory import PipelineScheduleInferenceConfig

from ._assembly import build_housekeeping, lay_foundation


class InferenceConfigurator:
    """Assembles an :class:`Inference` job (forward-only schedule; reference ``run/inference.py:55-200``)."""

    def __init__(self, mesh: DeviceMeshParameters, parameters: InferenceConfig, task_provider: control.InferenceTaskProvider,
                 model_provider: control.ModelProvider, data_provider: control.DatasetProvider):
        self._mesh, self._parameters = mesh, parameters
        self._task_provider, self._model_provider, self._data_provider = task_provider, model_provider, data_provider

    def _build_state(self) -> InferenceJobState:
        cfg = self._parameters
        pipelining = PipeliningConfig(schedule=PipelineScheduleInferenceConfig())  # forward-only program
        base = lay_foundation(self._mesh, cfg, lambda ctx: self._task_provider(control.InferenceTaskProviderContext(dist_context=ctx)),
                              self._model_provider, cfg.batching, pipelining, events.EVENT_INFERENCE_CONFIG_STARTED)
        ctx, bus, maths, task = base.ctx, base.bus, base.maths, base.task

        loader = parts.DataLoaderFactory(dist_context=ctx, provider=self._data_provider, config_data_loading=cfg.data_loading,
                                         batch_maths=maths).build_dataloader_for_infer_job()
        bus.trigger(events.EVENT_INFERENCE_DATA_LOADER_READY, shared_events.EventDataLoaderReadyContext(data_loader=loader))
        stepper = parts.Stepper(initial_step=0, total_steps=len(loader))

        pipeline_state = PipelineStateHandler(sharding_spec={}, num_shards=maths.num_microbatches_pipelining)
        schedule, modules = parts.ModelStageFactory(
            model_provider=self._model_provider, dist_context=ctx, batch_maths=maths, config_model=cfg.model_stage_factory,
            config_pipelining=pipelining, pipeline_callback=parts.InferenceProcessor(state=pipeline_state, task=task),
        ).build_pipeline_and_modules()
        bus.trigger(events.EVENT_INFERENCE_MODEL_STAGES_READY, shared_events.EventModelStagesReadyContext(modules=modules.modules))

        house = build_housekeeping(ctx, cfg, stepper, run_name=None)
        return InferenceJobState(
            dist_context=ctx, event_bus=bus, task=task, batch_maths=maths, data_loader=loader, stepper=stepper, tracked_modules=modules,
            task_operator=parts.InferenceTaskOperator(dist_context=ctx, task=task, pipeline=schedule, pipeline_state=pipeline_state),
            garbage_collector=house.gc, checkpointer=house.checkpointer, profiler=house.profiler, timeout_manager=base.timeout,
        )

    def configure(self) -> "Inference":
        return Inference(self._build_state())


class Inference:
    """Batched forward loop over a dataset under ``torch.inference_mode`` with resumable progress."""

    def __init__(self, state: InferenceJobState):
        self._state = state

    @property
    def state(self) -> InferenceJobState:
        return self._state

    def infer(self) -> None:
        s = self._state
        with torch.inference_mode():
            for module in s.tracked_modules.modules:
                module.eval()
            s.dist_context.wait_world()
            s.dist_context.logger.info("Trying to load last checkpoint before doing anything else")
            s.checkpointer.load_last_checkpoint(s)
            if s.stepper.current_step >= s.stepper.total_steps:
                s.dist_context.logger.info("Already ran, will do nothing")
                return
            s.dist_context.wait_world()
            step_ctx = shared_events.EventStepContext(stepper=s.stepper)
            with (
                tqdm(desc="Inference", total=s.stepper.total_steps, initial=s.stepper.current_step,
                     disable=not s.dist_context.is_local_main_process) as bar,
                s.garbage_collector as gc,
                s.profiler.open() as profiler,
            ):
                s.event_bus.trigger(events.EVENT_INFERENCE_READY, events.EventInferenceReadyContext())
                for batch_group in s.data_loader:
                    s.event_bus.trigger(events.EVENT_INFERENCE_STEP_PRE, step_ctx)
                    with s.event_bus.bounded(events.EVENT_INFERENCE_FORWARD_PRE, events.EVENT_INFERENCE_FORWARD_POST, step_ctx):
                        for batch in batch_group:
                            s.task_operator.forward(batch)
                    gc.collect_periodic()
                    if profiler:
                        profiler.step()
                    s.timeout_manager.set_periodic()
                    s.event_bus.trigger(events.EVENT_INFERENCE_STEP_POST, step_ctx)
                    s.stepper.step()
                    s.checkpointer.checkpoint_if_needed(s)
                    bar.update()
                s.checkpointer.wait_pending()
                s.task.finalize(control.FinalizeContext())
                s.event_bus.trigger(events.EVENT_INFERENCE_FINISHED, events.EventInferenceFinishedContext())
