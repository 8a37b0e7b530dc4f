from pathlib import Path
from typing import Literal

from pydantic import BaseModel

from d9d_b200.pipelining.factory import AnyPipelineScheduleConfig
from d9d_b200.tracker import AnyTrackerConfig, RunConfig

from .types import StepActionPeriod


class BatchingConfig(BaseModel):
    global_batch_size: int  # over all replicas and accumulation steps
    microbatch_size: int  # one forward on one device


class DeterminismConfig(BaseModel):
    base_seed: int


class PipeliningConfig(BaseModel):
    schedule: AnyPipelineScheduleConfig


class GarbageCollectionConfig(BaseModel):
    period_steps: StepActionPeriod


class DataLoadingConfig(BaseModel):
    num_workers: int
    pin_memory: bool
    persistent_workers: bool


class CheckpointingConfig(BaseModel):
    save_dir: Path
    period_steps: StepActionPeriod
    num_to_keep: int | None
    # stage the state in host memory and write it from a background thread while training continues (the next save,
    # a load and the end of the job wait for the write in flight)
    async_save: bool = False


class Fp8LinearConfig(BaseModel):
    """Run the selected dense linear layers in fp8 (e4m3, dynamic row / column scaling for forward, input-gradient and
    weight-gradient GEMMs - ``d9d_b200.kernel.fp8.Fp8Linear``).  Net-new: the reference has no fp8 path.

    ``include`` / ``exclude`` are regular expressions searched in the module names of a model stage; only plain linear
    layers whose feature sizes are multiples of 16 are converted (LoRA wrappers, grouped expert weights, norms are not)."""

    include: str = r".*"
    exclude: str | None = r"lm_head|cls_head|embedding_head|router|\.gate$"
    recipe: Literal["rowwise", "mx"] = "rowwise"  # "mx": MXFP8 block scaling (feature sizes and token count multiples of 128)


class ModelStageFactoryConfig(BaseModel):
    source_checkpoint: Path | None
    checkpoint_only_trainable_parameters: bool
    fp8_linear: Fp8LinearConfig | None = None


class GradientClippingConfig(BaseModel):
    max_norm: float | None
    log_total_steps: StepActionPeriod


class ProfilingConfig(BaseModel):
    enabled: bool
    traces_dir: Path
    period_steps: int
    warmup_steps: int
    active_steps: int


class JobLoggerConfig(BaseModel):
    period_steps: StepActionPeriod
    tracker: AnyTrackerConfig


class GradientManagerConfig(BaseModel):
    grad_dtype: str | None  # None => parameter dtype
    bucket_size_mb: int
    # fold the 1/sum(weights) scaling and the clip coefficient into the optimizer's gradient-scale input instead of
    # rewriting every gradient twice (only when every optimizer accepts a device ``grad_scale``, e.g. StochasticAdamW)
    fold_scaling_into_optimizer: bool = True


class TimeoutConfig(BaseModel):
    init_timeout: int = 10000
    step_timeout: int = 100


class TrainerConfig(BaseModel):
    run: RunConfig
    batching: BatchingConfig
    data_loading: DataLoadingConfig
    logging: JobLoggerConfig
    pipelining: PipeliningConfig
    model_stage_factory: ModelStageFactoryConfig
    determinism: DeterminismConfig
    gc: GarbageCollectionConfig
    checkpointing: CheckpointingConfig
    gradient_clipping: GradientClippingConfig
    profiling: ProfilingConfig | None
    gradient_manager: GradientManagerConfig
    timeout: TimeoutConfig = TimeoutConfig()


class InferenceConfig(BaseModel):
    batching: BatchingConfig
    data_loading: DataLoadingConfig
    model_stage_factory: ModelStageFactoryConfig
    determinism: DeterminismConfig
    gc: GarbageCollectionConfig
    checkpointing: CheckpointingConfig
    profiling: ProfilingConfig | None
    timeout: TimeoutConfig = TimeoutConfig()
