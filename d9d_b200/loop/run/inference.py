from __future__ import annotations

import torch
from tqdm import tqdm

from d9d_b200.core.dist_context import DeviceMeshParameters
from d9d_b200.internals.determinism import set_seeds
from d9d_b200.internals.pipeline_state import PipelineStateHandler
from d9d_b200.loop.component import (
    BatchMaths,
    DataLoaderFactory,
    InferenceProcessor,
    InferenceTaskOperator,
    JobProfiler,
    ManualGarbageCollector,
    ModelStageFactory,
    StateCheckpointer,
    Stepper,
    TimeoutManager,
)
from d9d_b200.loop.config import InferenceConfig, PipeliningConfig
from d9d_b200.loop.control import (
    DatasetProvider,
    FinalizeContext,
    InferenceTaskProvider,
    InferenceTaskProviderContext,
    ModelProvider,
    RegisterModelEventsContext,
    RegisterTaskEventsContext,
)
from d9d_b200.loop.event import EventBus
from d9d_b200.loop.event.catalogue.common import (
    EventConfigurationStartedContext,
    EventDataLoaderReadyContext,
    EventModelStagesReadyContext,
    EventStepContext,
)
from d9d_b200.loop.event.catalogue.inference import (
    EVENT_INFERENCE_CONFIG_STARTED,
    EVENT_INFERENCE_DATA_LOADER_READY,
    EVENT_INFERENCE_FINISHED,
    EVENT_INFERENCE_FORWARD_POST,
    EVENT_INFERENCE_FORWARD_PRE,
    EVENT_INFERENCE_MODEL_STAGES_READY,
    EVENT_INFERENCE_READY,
    EVENT_INFERENCE_STEP_POST,
    EVENT_INFERENCE_STEP_PRE,
    EventInferenceFinishedContext,
    EventInferenceReadyContext,
)
from d9d_b200.loop.state import InferenceJobState
from d9d_b200.pipelining.factory import PipelineScheduleInferenceConfig


class InferenceConfigurator:
    """Assembles an :class:`Inference` job (forward-only schedule; reference ``run/inference.py:55-200``)."""

    def __init__(self, mesh: DeviceMeshParameters, parameters: InferenceConfig, task_provider: InferenceTaskProvider,
                 model_provider: ModelProvider, data_provider: DatasetProvider):
        self._mesh, self._parameters = mesh, parameters
        self._task_provider, self._model_provider, self._data_provider = task_provider, model_provider, data_provider

    def _build_state(self) -> InferenceJobState:
        cfg = self._parameters
        ctx = self._mesh.build()
        pipelining = PipeliningConfig(schedule=PipelineScheduleInferenceConfig())
        set_seeds(ctx, seed=cfg.determinism.base_seed)
        timeout = TimeoutManager(dist_context=ctx, config=cfg.timeout)
        timeout.set_init()
        task = self._task_provider(InferenceTaskProviderContext(dist_context=ctx))
        bus = EventBus()
        self._model_provider.register_events(RegisterModelEventsContext(dist_context=ctx, event_bus=bus))
        task.register_events(RegisterTaskEventsContext(dist_context=ctx, event_bus=bus))
        bus.trigger(EVENT_INFERENCE_CONFIG_STARTED, EventConfigurationStartedContext(dist_context=ctx))

        maths = BatchMaths(dist_context=ctx, config_batching=cfg.batching, config_pipelining=pipelining)
        loader = DataLoaderFactory(dist_context=ctx, provider=self._data_provider, config_data_loading=cfg.data_loading,
                                   batch_maths=maths).build_dataloader_for_infer_job()
        bus.trigger(EVENT_INFERENCE_DATA_LOADER_READY, EventDataLoaderReadyContext(data_loader=loader))
        stepper = Stepper(initial_step=0, total_steps=len(loader))
        pipeline_state = PipelineStateHandler(sharding_spec={}, num_shards=maths.num_microbatches_pipelining)
        processor = InferenceProcessor(state=pipeline_state, task=task)
        schedule, modules = ModelStageFactory(model_provider=self._model_provider, dist_context=ctx, batch_maths=maths,
                                              config_model=cfg.model_stage_factory, config_pipelining=pipelining,
                                              pipeline_callback=processor).build_pipeline_and_modules()
        bus.trigger(EVENT_INFERENCE_MODEL_STAGES_READY, EventModelStagesReadyContext(modules=modules.modules))
        operator = InferenceTaskOperator(dist_context=ctx, task=task, pipeline=schedule, pipeline_state=pipeline_state)
        gc = ManualGarbageCollector(dist_ctx=ctx, config=cfg.gc, step=stepper)
        checkpointer = StateCheckpointer(dist_context=ctx, stepper=stepper, config=cfg.checkpointing, gc=gc, run_name=None)
        profiler = JobProfiler(dist_context=ctx, stepper=stepper, config=cfg.profiling)
        return InferenceJobState(dist_context=ctx, data_loader=loader, stepper=stepper, tracked_modules=modules, garbage_collector=gc,
                                 batch_maths=maths, checkpointer=checkpointer, task=task, profiler=profiler, timeout_manager=timeout,
                                 task_operator=operator, event_bus=bus)

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
            step_ctx = EventStepContext(stepper=s.stepper)
            with (
                tqdm(desc="Inference", total=s.stepper.total_steps, initial=s.stepper.current_step,
                     disable=not s.dist_context.is_local_main_process) as bar,
                s.garbage_collector as gc,
                s.profiler.open() as profiler,
            ):
                s.event_bus.trigger(EVENT_INFERENCE_READY, EventInferenceReadyContext())
                for batch_group in s.data_loader:
                    s.event_bus.trigger(EVENT_INFERENCE_STEP_PRE, step_ctx)
                    with s.event_bus.bounded(EVENT_INFERENCE_FORWARD_PRE, EVENT_INFERENCE_FORWARD_POST, step_ctx):
                        for batch in batch_group:
                            s.task_operator.forward(batch)
                    gc.collect_periodic()
                    if profiler:
                        profiler.step()
                    s.timeout_manager.set_periodic()
                    s.event_bus.trigger(EVENT_INFERENCE_STEP_POST, step_ctx)
                    s.stepper.step()
                    s.checkpointer.checkpoint_if_needed(s)
                    bar.update()
                s.task.finalize(FinalizeContext())
                s.event_bus.trigger(EVENT_INFERENCE_FINISHED, EventInferenceFinishedContext())
