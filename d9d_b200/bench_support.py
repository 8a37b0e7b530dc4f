"""Training-step driver used by ``bench.py`` and ``__graft_entry__.smoke``.

One optimizer step = ``accum`` microbatches of forward+backward (loss = per-token mean weighted by token count,
exactly the reference's SFT task contract), gradients accumulated in a flat fp32 arena that every ``param.grad``
aliases, data-parallel SUM all-reduce of the arena, device-side ``1/sum(weight)`` scaling + global-norm clipping
folded into the fused stochastic-rounding AdamW launch.  No host synchronisation inside a step.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist

from d9d_b200.kernel._native import native_launch_count, native_ops
from d9d_b200.optim.stochastic import StochasticAdamW

LM_IGNORE_INDEX = -100


class TrainStepRunner:
    def __init__(self, args, device: torch.device, world: int, build_model, lr: float = 2.5e-4, max_norm: float = 5.0):
        self.args = args
        self.device = device
        self.world = world
        if world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
        torch.manual_seed(1337)  # identical init on every replica
        self.model = build_model(args, device)
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        self.grad_scale = torch.ones(1, dtype=torch.float32, device=device)
        self.max_norm = max_norm
        self.nvlink = world > 1 and getattr(args, "dp_impl", "nvlink") == "nvlink"
        if self.nvlink:
            # gradients / parameters live in symmetric NVLink arenas; reduce-scatter + AdamW + all-gather are two
            # peer-memory kernels inside opt.step() (no NCCL all-reduce of the gradients)
            from d9d_b200.optim.nvlink import NvlinkShardedAdamW

            try:
                self.opt = NvlinkShardedAdamW(self.params, dist.group.WORLD, lr=lr, state_dtype=torch.bfloat16, max_norm=max_norm)
            except Exception as exc:  # e.g. no peer access between the visible GPUs: every rank fails the same way
                import sys

                print(f"[bench] NVLink symmetric memory unavailable ({type(exc).__name__}: {exc}); falling back to NCCL all-reduce",
                      file=sys.stderr, flush=True)
                self.nvlink = False
                args.dp_impl = "nccl"
        if self.nvlink:
            self.opt.grad_scale = self.grad_scale
            self.opt.set_required_accumulations(args.accum)  # reduce finished chunks while backward is still running
            self.grad_arena = self.opt.grad_arena.buffer
        else:
            # flat fp32 gradient arena; every param.grad is a view (accumulated in place by autograd)
            total = sum((p.numel() + 63) // 64 * 64 for p in self.params)
            self.grad_arena = torch.zeros(total, dtype=torch.float32, device=device)
            off = 0
            for p in self.params:
                p.grad_dtype = torch.float32
                p.grad = self.grad_arena[off : off + p.numel()].view_as(p)
                p._d9d_fused_wgrad = True  # wgrad GEMMs accumulate into the arena in their epilogue
                off += (p.numel() + 63) // 64 * 64
            self.opt = StochasticAdamW(self.params, lr=lr, state_dtype=torch.bfloat16)
            self.opt.grad_scale = self.grad_scale
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        self.weight_sum = torch.zeros(1, dtype=torch.float32, device=device)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=device)
        self.comm_stream = torch.cuda.Stream(device=device) if world > 1 else None
        self.pos = torch.arange(args.seq_len, device=device)[None].expand(args.microbatch, -1).contiguous()

    # ---------------------------------------------------------------- data
    def synthetic_batch(self, vocab: int, gen: torch.Generator) -> dict[str, torch.Tensor]:
        tok = torch.randint(0, vocab, (self.args.microbatch, self.args.seq_len + 1), device=self.device, generator=gen)
        return {"input_ids": tok[:, :-1].contiguous(), "labels": tok[:, 1:].contiguous()}

    def synthetic_host_batch(self, vocab: int, seed: int) -> dict[str, torch.Tensor]:
        g = torch.Generator().manual_seed(seed)
        tok = torch.randint(0, vocab, (self.args.microbatch, self.args.seq_len + 1), generator=g)
        return {"input_ids": tok[:, :-1].contiguous().pin_memory(), "labels": tok[:, 1:].contiguous().pin_memory()}

    # ---------------------------------------------------------------- one optimizer step
    def _forward_backward(self, batch: dict[str, torch.Tensor]) -> None:
        labels = batch["labels"]
        out = self.model(input_ids=batch["input_ids"], position_ids=self.pos, labels=labels)
        n_tok = (labels != LM_IGNORE_INDEX).sum()
        loss = out["logps"].sum() / n_tok
        weight = n_tok / 1000.0
        (loss * weight).backward()
        self.weight_sum += weight.detach()
        self.loss_sum += (loss.detach() * weight.detach()).float()

    def step(self, batches: list[dict[str, torch.Tensor]]) -> torch.Tensor:
        for b in batches:
            self._forward_backward(b)
        ops = native_ops()
        if self.world > 1:
            if not self.nvlink:
                dist.all_reduce(self.grad_arena)
            stats = torch.cat([self.weight_sum, self.loss_sum])
            dist.all_reduce(stats)
            self.weight_sum, self.loss_sum = stats[:1].clone(), stats[1:].clone()
        inv_w = 1.0 / self.weight_sum
        if self.nvlink:
            self.grad_scale.copy_(inv_w)  # reduction, clipping, update, broadcast and zeroing happen inside step()
            self.opt.step()
        else:
            # global grad norm of the *scaled* gradient, clip coefficient, all on the device
            self.sumsq.zero_()
            ops.sumsq_accumulate_(self.grad_arena, self.sumsq)
            norm = self.sumsq.sqrt() * inv_w
            clip = torch.clamp(self.max_norm / (norm + 1e-6), max=1.0)
            self.grad_scale.copy_(inv_w * clip)
            self.opt.step()
            self.grad_arena.zero_()
        loss = (self.loss_sum / self.weight_sum).clone()
        self.weight_sum.zero_()
        self.loss_sum.zero_()
        return loss

    def step_from_host(self, host_batches: list[dict[str, torch.Tensor]]) -> tuple[float, int, int]:
        h2d = 0
        dev_batches = []
        for hb in host_batches:
            db = {k: v.to(self.device, non_blocking=True) for k, v in hb.items()}
            h2d += sum(v.numel() * v.element_size() for v in hb.values())
            dev_batches.append(db)
        loss = self.step(dev_batches)
        loss_host = loss.to("cpu")  # device -> host read of the step's result (synchronises this step)
        return float(loss_host), h2d, loss_host.numel() * loss_host.element_size()

    def launch_count(self) -> int:
        return native_launch_count()


# ------------------------------------------------------------------------------------------------------------------
# End-to-end measurement through the public training API (TrainingConfigurator -> Trainer.train())
# ------------------------------------------------------------------------------------------------------------------
def trainer_config_for_bench(args, world: int, total_steps: int, workdir: str) -> dict:
    """The reference example's trainer section (``example/qwen3_moe/pretrain.json``) with checkpointing/tracking off."""
    return {
        "run": {"name": "bench", "description": None, "hparams": {}},
        "batching": {"global_batch_size": args.accum * args.microbatch * world, "microbatch_size": args.microbatch},
        "data_loading": {"num_workers": 0, "pin_memory": True, "persistent_workers": False},
        "logging": {"period_steps": 10_000, "tracker": {"provider": "null"}},
        "pipelining": {"schedule": ({"schedule": "looped_bfs", "num_stages_per_rank": 2} if getattr(args, "layout", "dp") == "example"
                                     else {"schedule": "gpipe"})},
        "model_stage_factory": {"source_checkpoint": None, "checkpoint_only_trainable_parameters": False},
        "determinism": {"base_seed": 1337},
        "gc": {"period_steps": 10_000},
        "checkpointing": {"save_dir": workdir, "period_steps": "disable", "num_to_keep": 1},
        "gradient_clipping": {"max_norm": 5.0, "log_total_steps": 10_000},
        "profiling": None,
        "gradient_manager": {"grad_dtype": "float32", "bucket_size_mb": 32},
        "timeout": {"init_timeout": 10_000, "step_timeout": 600},
    }


class TrainerEndToEnd:
    """Runs ``warmup + steps`` optimizer steps through ``Trainer.train()`` and times the last ``steps`` of them.

    Every step's batches come out of the framework's ``StatefulDataLoader`` in pinned host memory and are copied to
    the device inside the loop; the step's loss is copied back to the host by the job logger.  Timing uses CUDA
    events recorded from ``EVENT_TRAIN_STEP_PRE`` / ``EVENT_TRAIN_STEP_POST`` subscribers.
    """

    def __init__(self, args, world: int, model_params, vocab: int, workdir: str, layout: str = "dp"):
        from d9d_b200.core.dist_context import DeviceMeshParameters
        from d9d_b200.loop.auto import AutoLRSchedulerProvider, AutoOptimizerProvider
        from d9d_b200.loop.auto.auto_lr_scheduler import PiecewiseConfig
        from d9d_b200.loop.auto.auto_optimizer import NvlinkShardedAdamWOptimizerConfig, StochasticAdamWOptimizerConfig
        from d9d_b200.loop.config import TrainerConfig
        from d9d_b200.loop.run import TrainingConfigurator
        from d9d_b200.recipes import (
            CausalLMTask,
            Qwen3MoEModelProvider,
            Qwen3MoEModelProviderConfig,
            SyntheticDataConfig,
            SyntheticDataProvider,
        )

        self.args, self.world = args, world
        self.total_steps = args.warmup + args.steps
        global_batch = args.accum * args.microbatch * world
        lr_cfg = PiecewiseConfig.model_validate({"name": "piecewise", "scheduler": {"initial_multiplier": 1.0, "phases": [
            {"mode": "rest", "target_multiplier": 1.0, "curve": {"type": "linear"}}]}})
        self.layout = layout
        nvlink_dp = layout == "dp" and world > 1 and getattr(args, "dp_impl", "nvlink") == "nvlink"
        if layout == "ep":
            mesh = DeviceMeshParameters(data_parallel_replicate=world, expert_parallel=world)
        elif layout == "example":  # the reference example's own layout (example/qwen3_moe/pretrain.json:2-10), 8 GPUs
            if world != 8:
                raise ValueError("--layout example is the reference's PP4 x DP2 x EP2 layout: it needs 8 GPUs")
            mesh = DeviceMeshParameters(pipeline_parallel=4, data_parallel_replicate=2, expert_parallel=2)
        else:
            mesh = DeviceMeshParameters(data_parallel_replicate=world)
        self.trainer = TrainingConfigurator(
            mesh=mesh,
            parameters=TrainerConfig.model_validate(trainer_config_for_bench(args, world, self.total_steps, workdir)),
            task_provider=lambda ctx: CausalLMTask(),
            model_provider=Qwen3MoEModelProvider(Qwen3MoEModelProviderConfig(model=model_params, checkpointing=bool(getattr(args, "checkpointing", False)))),
            data_provider=SyntheticDataProvider(SyntheticDataConfig(
                num_samples=global_batch * self.total_steps, seq_len=args.seq_len, vocab_size=vocab, seed=5)),
            optimizer_provider=AutoOptimizerProvider(
                NvlinkShardedAdamWOptimizerConfig(lr=2.5e-4, state_dtype="bfloat16") if nvlink_dp
                else StochasticAdamWOptimizerConfig(lr=2.5e-4, state_dtype="bfloat16")),
            lr_scheduler_provider=AutoLRSchedulerProvider(lr_cfg),
        ).configure()

    def run(self) -> dict:
        from d9d_b200.loop.event.catalogue.train import EVENT_TRAIN_STEP_POST, EVENT_TRAIN_STEP_PRE

        args, state = self.args, self.trainer.state
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = {}

        prof_path = os.environ.get("D9D_BENCH_PROFILE")  # attribution only: per-kernel totals of the LAST step (rank 0)
        prof_box: dict = {}

        def pre(ctx) -> None:
            if prof_path and state.dist_context.is_main_process and ctx.stepper.current_step == self.total_steps - 1:
                from torch.profiler import ProfilerActivity, profile

                torch.cuda.synchronize()
                prof_box["p"] = profile(activities=[ProfilerActivity.CUDA])
                prof_box["p"].__enter__()
            if ctx.stepper.current_step == args.warmup:
                state.dist_context.wait_world()
                launches["before"] = native_launch_count()
                start.record()

        def post(ctx) -> None:
            if ctx.stepper.current_step == self.total_steps - 1:
                end.record()
                if "p" in prof_box:
                    torch.cuda.synchronize()
                    prof_box["p"].__exit__(None, None, None)
                    totals: dict[str, list] = {}
                    for ev in prof_box["p"].events():
                        if ev.device_type == torch.autograd.DeviceType.CUDA:
                            t = totals.setdefault(ev.name, [0, 0.0])
                            t[0] += 1
                            t[1] += ev.device_time
                    rows = sorted(totals.items(), key=lambda kv: -kv[1][1])
                    with open(prof_path, "w") as f:
                        import json as _json

                        _json.dump([{"name": k[:160], "calls": v[0], "total_us": v[1]} for k, v in rows], f, indent=1)
                state.dist_context.wait_world()
                launches["after"] = native_launch_count()

        state.event_bus.subscribe(EVENT_TRAIN_STEP_PRE, pre)
        state.event_bus.subscribe(EVENT_TRAIN_STEP_POST, post)
        self.trainer.train()
        torch.cuda.synchronize()
        ms = torch.tensor([start.elapsed_time(end)], device=state.dist_context.current_device)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # bytes moved per optimizer step and rank: input_ids + labels + position_ids (int64) up, one fp32 loss down
        h2d = args.accum * 3 * args.microbatch * args.seq_len * 8
        overflow = None
        if self.layout in ("ep", "example"):
            from d9d_b200.module.block.moe.communications.nvlink import NvlinkExpertParallelCommunicationHandler as H

            overflow = any(h.overflowed for h in H.instances)
        return {"ms_per_step": ms.item() / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "final_loss": state.logger.last_loss, "launches": launches.get("after", 0) - launches.get("before", 0),
                "ep_overflow": overflow}
