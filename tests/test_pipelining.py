"""Pipeline programs (pure python), split backward (single process) and end-to-end PP on gloo (4 processes)."""

import pytest
import torch
from torch import nn

from d9d_b200.core.autograd import GLOBAL_GRAD_CONTEXT, GradDirection
from d9d_b200.pipelining.api import PipelineStageInfo, distribute_layers_for_pipeline_stage
from d9d_b200.pipelining.infra.action import ActionKind, flatten
from d9d_b200.pipelining.infra.programs import (
    DualPipeVPipelineProgramBuilder,
    Interleaved1F1BPipelineProgramBuilder,
    LoopedBFSPipelineProgramBuilder,
    ZeroBubbleVPipelineProgramBuilder,
)
from d9d_b200.pipelining.infra.stage.split_backward import backward_full, backward_input, backward_weight
from tests.dist_utils import run_distributed

BUILDERS = {
    "gpipe": lambda: LoopedBFSPipelineProgramBuilder(1),
    "inference": lambda: LoopedBFSPipelineProgramBuilder(1, inference_mode=True),
    "looped_bfs2": lambda: LoopedBFSPipelineProgramBuilder(2),
    "1f1b": lambda: Interleaved1F1BPipelineProgramBuilder(1),
    "1f1b_interleaved": lambda: Interleaved1F1BPipelineProgramBuilder(2),
    "zb1p": lambda: Interleaved1F1BPipelineProgramBuilder(2, enable_zero_bubble=True),
    "zbv": ZeroBubbleVPipelineProgramBuilder,
    "dualpipev": DualPipeVPipelineProgramBuilder,
}


@pytest.mark.parametrize("name", list(BUILDERS))
@pytest.mark.parametrize("pp,mb", [(2, 1), (2, 4), (4, 8), (4, 5), (8, 32), (3, 7)])
def test_programs_complete_and_deadlock_free(name, pp, mb):
    builder = BUILDERS[name]()
    if name == "dualpipev" and mb < 2 * pp:
        with pytest.raises(ValueError):
            builder.compose(mb, pp)
        return
    program = builder.compose(mb, pp)  # compose() replays the program and raises on deadlock / missing work
    stages = builder.num_stages_per_rank * pp
    acts = [a for r in program.values() for slot in r for a in flatten(slot)]
    assert sum(a.kind == ActionKind.FORWARD for a in acts) == stages * mb
    sends = sum(a.kind == ActionKind.SEND_F for a in acts)
    recvs = sum(a.kind == ActionKind.RECV_F for a in acts)
    assert sends == recvs
    if builder.has_backward:
        full = sum(a.kind == ActionKind.BACKWARD_FULL for a in acts)
        inp = sum(a.kind == ActionKind.BACKWARD_INPUT for a in acts)
        wgt = sum(a.kind == ActionKind.BACKWARD_WEIGHT for a in acts)
        assert full + inp == stages * mb and inp == wgt
    else:
        assert not any(a.has_backward_work for a in acts)


def test_1f1b_matches_textbook_order():
    program = Interleaved1F1BPipelineProgramBuilder(1).compose(8, 4)
    compute = [str(a) for a in program[3] if a.is_compute]
    assert compute[:6] == ["3F0", "3B0", "3F1", "3B1", "3F2", "3B2"]
    compute0 = [str(a) for a in program[0] if a.is_compute]
    assert compute0[:5] == ["0F0", "0F1", "0F2", "0F3", "0B0"]  # warm-up depth P on the first rank


def test_layer_distribution():
    assert distribute_layers_for_pipeline_stage(16, 0, 1, PipelineStageInfo(0, 8)) == (0, 3)
    assert distribute_layers_for_pipeline_stage(16, 0, 1, PipelineStageInfo(7, 8)) == (15, 16)
    spans = [distribute_layers_for_pipeline_stage(10, 1, 1, PipelineStageInfo(i, 4)) for i in range(4)]
    assert spans == [(0, 2), (2, 5), (5, 8), (8, 10)]
    with pytest.raises(ValueError):
        distribute_layers_for_pipeline_stage(2, 2, 2, PipelineStageInfo(0, 4))


# ----------------------------------------------------------------------------- split backward
class _DirLinear(torch.autograd.Function):
    """Linear honouring GLOBAL_GRAD_CONTEXT (like the native ops do) and counting which grads it computed."""

    calls = {"dx": 0, "dw": 0}

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return x @ w.t()

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0] and GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.inputs):
            _DirLinear.calls["dx"] += 1
            dx = g @ w
        if ctx.needs_input_grad[1] and GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.weight):
            _DirLinear.calls["dw"] += 1
            dw = g.t() @ x
        return dx, dw


class _Toy(nn.Module):
    def __init__(self, d=8):
        super().__init__()
        self.w1 = nn.Parameter(torch.randn(d, d) * 0.3)
        self.w2 = nn.Parameter(torch.randn(d, d) * 0.3)
        self.norm = nn.LayerNorm(d)
        self.shared = nn.Parameter(torch.randn(d) * 0.1)
        self.frozen = nn.Parameter(torch.randn(d, d) * 0.3, requires_grad=False)

    def forward(self, x):
        h = _DirLinear.apply(x, self.w1) + self.shared
        h = self.norm(torch.tanh(h)) @ self.frozen
        return _DirLinear.apply(h, self.w2) * self.shared


def test_split_backward_matches_full():
    torch.manual_seed(0)
    m = _Toy()
    x = torch.randn(5, 8, requires_grad=True)
    gout = torch.randn(5, 8)

    y = m(x)
    (dx_ref,) = backward_full([y], [gout], [x])
    ref = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()

    hook_hits = {n: 0 for n, p in m.named_parameters() if p.requires_grad}
    for n, p in m.named_parameters():
        if p.requires_grad:
            p.register_post_accumulate_grad_hook(lambda _p, n=n: hook_hits.__setitem__(n, hook_hits[n] + 1))
    _DirLinear.calls.update(dx=0, dw=0)
    y = m(x)
    (dx,), deferred = backward_input([y], [gout], [x], [p for p in m.parameters() if p.requires_grad])
    assert _DirLinear.calls == {"dx": 2, "dw": 0}  # the input pass skipped every weight gradient
    assert all(p.grad is None for p in m.parameters())
    torch.testing.assert_close(dx, dx_ref)
    backward_weight(deferred)
    assert _DirLinear.calls["dw"] == 2
    for n, p in m.named_parameters():
        if p.requires_grad:
            torch.testing.assert_close(p.grad, ref[n])
    assert all(v == 1 for v in hook_hits.values()), hook_hits  # one accumulate event per parameter (also the shared one)
    assert m.frozen.grad is None


# ----------------------------------------------------------------------------- end-to-end PP on gloo
class _StageModel(nn.Module):
    """Toy pipeline stage: three matmuls through the direction-aware op; first stage embeds, last returns 'out'."""

    def __init__(self, stage: PipelineStageInfo, d=16, freeze_some=False):
        super().__init__()
        g = torch.Generator().manual_seed(1000 + stage.current_stage)
        self.ws = nn.ParameterList([nn.Parameter(torch.randn(d, d, generator=g) * 0.2) for _ in range(3)])
        if freeze_some:
            self.ws[1].requires_grad_(False)
        self._stage = stage
        self._d = d

    def forward(self, x=None, hidden=None, scale=None):
        h = x if hidden is None else hidden
        for w in self.ws:
            h = torch.tanh(_DirLinear.apply(h, w))
        if scale is not None:
            h = h * scale[:, None]
        return {"hidden": h}

    def infer_stage_inputs_from_pipeline_inputs(self, inputs, n_microbatches):
        x = inputs["x"]
        key = "x" if self._stage.is_current_stage_first else "hidden"
        return {key: torch.empty((x.shape[0] // n_microbatches, self._d), dtype=torch.float32, device=x.device)}

    def infer_stage_outputs_from_pipeline_inputs(self, inputs, n_microbatches):
        x = inputs["x"]
        return {"hidden": torch.empty((x.shape[0] // n_microbatches, self._d), dtype=torch.float32, device=x.device)}


def _pp_worker(rank, world, cases):
    """All schedules on one process group (spawning processes dominates the cost of these tests)."""
    from d9d_b200.core.dist_context import DeviceMeshParameters

    ctx = DeviceMeshParameters(pipeline_parallel=world).build()
    for cfg_json, n_mb in cases:
        try:
            _pp_case(ctx, cfg_json, n_mb, n_mb % 2 == 0 and "zero" in cfg_json)
        except AssertionError as exc:
            raise AssertionError(f"[{cfg_json} microbatches={n_mb}] {exc}") from exc
        ctx.wait_world()


def _pp_case(ctx, cfg_json, n_mb, freeze):
    from pydantic import TypeAdapter

    from d9d_b200.pipelining.factory import AnyPipelineScheduleConfig, build_schedule

    cfg = TypeAdapter(AnyPipelineScheduleConfig).validate_json(cfg_json)
    batch = n_mb * 2
    dev = ctx.current_device  # cpu (gloo) or this rank's GPU (NCCL p2p, tests/test_parallelism_gpu.py)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(batch, 16, generator=g).to(dev)
    scale = (torch.rand(batch, generator=g) + 0.5).to(dev)
    target = torch.randn(batch, 16, generator=g).to(dev)
    target_mb = target.chunk(n_mb)
    losses = {}

    def loss_fn(outputs, mb):
        loss = ((outputs["hidden"] - target_mb[mb]) ** 2).sum()
        losses[mb] = loss.detach()
        return loss

    results = {}

    def result_fn(outputs, mb):
        results[mb] = outputs["hidden"].detach().clone()

    hits = {}
    is_inference = cfg.schedule == "inference"
    info, modules = build_schedule(ctx, n_mb, cfg, lambda st: _StageModel(st, freeze_some=freeze).to(dev),
                                   result_fn if is_inference else loss_fn)
    for mi, mod in enumerate(modules):
        for n, p in mod.named_parameters():
            if p.requires_grad:
                hits[(mi, n)] = 0
                p.register_post_accumulate_grad_hook(lambda _p, k=(mi, n): hits.__setitem__(k, hits[k] + 1))
    for _ in range(2):  # two steps: caches must be drained between them
        for mod in modules:
            mod.zero_grad()
        for k in hits:
            hits[k] = 0
        info.schedule.configure_buffers({"x": x}, {"scale": scale}, None)
        info.schedule.step({"x": x}, {"scale": scale})
    if not is_inference:
        assert all(v == n_mb for v in hits.values()), hits

    # sequential reference on every rank
    num_stages = modules[0]._stage.num_stages
    ref_stages = [_StageModel(PipelineStageInfo(s, num_stages), freeze_some=freeze).to(dev) for s in range(num_stages)]
    h = x
    for s, st in enumerate(ref_stages):
        h = st(x=h, scale=scale)["hidden"] if s == 0 else st(hidden=h, scale=scale)["hidden"]
    if is_inference:
        if info.has_last_stage:  # every microbatch reached the result callback with the sequential model's output
            assert sorted(results) == list(range(n_mb))
            torch.testing.assert_close(torch.cat([results[i] for i in range(n_mb)]), h.detach(), rtol=1e-5, atol=1e-6)
        else:
            assert not results
        assert all(p.grad is None for mod in modules for p in mod.parameters())
        return
    ((h - target) ** 2).sum().backward()
    for mod in modules:
        ref = ref_stages[mod._stage.current_stage]
        for (n, p), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
            if p.requires_grad:
                torch.testing.assert_close(p.grad, q.grad, rtol=1e-4, atol=1e-5, msg=lambda m, n=n: f"{n}: {m}")
            else:
                assert p.grad is None
    if info.has_last_stage:
        assert len(losses) == n_mb


_PP_CASES = [
    ('{"schedule":"gpipe"}', 4),
    ('{"schedule":"inference"}', 3),
    ('{"schedule":"looped_bfs","num_stages_per_rank":2}', 4),
    ('{"schedule":"1f1b","num_stages_per_rank":1,"zero_bubble":false}', 8),
    ('{"schedule":"1f1b","num_stages_per_rank":2,"zero_bubble":false}', 4),
    ('{"schedule":"1f1b","num_stages_per_rank":2,"zero_bubble":true}', 8),
    ('{"schedule":"zero_bubble_v"}', 6),
    ('{"schedule":"dual_pipe_v"}', 8),
]


@pytest.mark.dist
def test_pipeline_end_to_end_gloo():
    """Every schedule on 4 pipeline ranks: gradients equal the sequential model, hooks fire once per microbatch, caches
    drain between steps, inference delivers every microbatch's output to the result callback."""
    run_distributed(_pp_worker, 4, _PP_CASES)
