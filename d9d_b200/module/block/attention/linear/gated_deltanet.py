from __future__ import annotations

import math
from typing import Annotated, Literal

import torch
import torch.nn.functional as F
from pydantic import BaseModel, Field
from torch import nn

from d9d_b200.kernel.linear_attn import causal_conv1d_silu, chunk_gated_delta_rule, mamba_decay_gate
from d9d_b200.kernel.swiglu import silu_mul
from d9d_b200.module.base import ModuleLateInit
from d9d_b200.module.block.linear import Linear
from d9d_b200.module.block.normalization import RMSNorm


class CausalShortDepthwiseConv1d(nn.Module, ModuleLateInit):
    """Left-padded depthwise convolution over the sequence followed by SiLU (weight ``[channels, kernel]``).

    Parity: reference ``attention/linear/gated_deltanet.py:17-73``.
    """

    def __init__(self, hidden_size: int, kernel_size: int) -> None:
        super().__init__()
        self._kernel_size = kernel_size
        self.weight = nn.Parameter(torch.empty(hidden_size, kernel_size))

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        if mask is not None:
            x = x * mask.unsqueeze(-1)
        return causal_conv1d_silu(x, self.weight)

    def reset_parameters(self) -> None:
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class LogSigmoidDecayGate(nn.Module, ModuleLateInit):
    """``logsigmoid(proj(x)) / normalizer`` (GLA / DeltaNet / HGRN-2 style); values in ``(-inf, 0]``."""

    def __init__(self, hidden_size: int, num_heads: int, normalizer: float = 16.0) -> None:
        super().__init__()
        self.proj = Linear(hidden_size, num_heads, bias=False)
        self._normalizer = normalizer

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.logsigmoid(self.proj(x).float()) / self._normalizer

    def reset_parameters(self) -> None:
        self.proj.reset_parameters()


class MambaDecayGate(nn.Module, ModuleLateInit):
    """``-exp(A_log) * softplus(proj(x) + dt_bias)`` (Mamba-2 / Qwen3-Next style) with the usual ``dt`` initialisation.

    Parity: reference ``gated_deltanet.py:112-186``.
    """

    def __init__(self, hidden_size: int, num_heads: int, normalizer: float = 16.0, dt_min: float = 0.001, dt_max: float = 0.1,
                 dt_init_floor: float = 1e-4) -> None:
        super().__init__()
        self.proj = Linear(hidden_size, num_heads, bias=False)
        self.A_log = nn.Parameter(torch.empty(num_heads, dtype=torch.float32))
        self.dt_bias = nn.Parameter(torch.empty(num_heads, dtype=torch.float32))
        self._num_heads = num_heads
        self._normalizer = normalizer
        self._dt_min, self._dt_max, self._dt_init_floor = dt_min, dt_max, dt_init_floor
        self.reset_parameters()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return mamba_decay_gate(self.proj(x), self.A_log, self.dt_bias)

    def reset_parameters(self) -> None:
        self.proj.reset_parameters()
        if self.A_log.is_meta:
            return
        with torch.no_grad():
            self.A_log.copy_(torch.empty_like(self.A_log).uniform_(0.0, self._normalizer).clamp_min(1e-6).log())
            span = math.log(self._dt_max) - math.log(self._dt_min)
            dt = torch.exp(torch.rand(self._num_heads, device=self.dt_bias.device) * span + math.log(self._dt_min))
            dt = dt.clamp(min=self._dt_init_floor)
            self.dt_bias.copy_(dt + torch.log(-torch.expm1(-dt)))  # inverse softplus


class MambaDecayGateParameters(BaseModel):
    type: Literal["mamba"] = "mamba"
    normalizer: float
    dt_min: float
    dt_max: float
    dt_init_floor: float


class LogSigmoidDecayGateParameters(BaseModel):
    type: Literal["logsigmoid"] = "logsigmoid"
    normalizer: float


AnyDecayGateParameters = Annotated[MambaDecayGateParameters | LogSigmoidDecayGateParameters, Field(discriminator="type")]


def _build_decay_gate(config: MambaDecayGateParameters | LogSigmoidDecayGateParameters, hidden_size: int, num_heads: int) -> nn.Module:
    if isinstance(config, MambaDecayGateParameters):
        return MambaDecayGate(hidden_size, num_heads, normalizer=config.normalizer, dt_min=config.dt_min, dt_max=config.dt_max,
                              dt_init_floor=config.dt_init_floor)
    if isinstance(config, LogSigmoidDecayGateParameters):
        return LogSigmoidDecayGate(hidden_size, num_heads, normalizer=config.normalizer)
    raise ValueError(f"Unknown decay gate config type: {type(config)}")


class GatedDeltaNet(nn.Module, ModuleLateInit):
    """Gated DeltaNet token mixer: fused q/k/v projection → causal short conv + SiLU → data-dependent decay and write
    strength → (GQA-expanded) chunked gated delta rule → per-head RMSNorm → SiLU output gate → output projection.

    Parity: reference ``attention/linear/gated_deltanet.py:246-386`` (which calls the ``fla`` Triton kernels).
    """

    def __init__(self, hidden_size: int, num_query_key_heads: int, num_value_heads: int, head_qk_dim: int, head_v_dim: int,
                 norm_eps: float, conv_size: int, decay_gate: MambaDecayGateParameters | LogSigmoidDecayGateParameters,
                 use_qk_l2norm: bool = True) -> None:
        super().__init__()
        if num_value_heads % num_query_key_heads != 0:
            raise ValueError(f"num_value_heads ({num_value_heads}) must be divisible by num_query_key_heads ({num_query_key_heads}).")
        self._num_qk_heads, self._num_v_heads = num_query_key_heads, num_value_heads
        self._groups = num_value_heads // num_query_key_heads
        self._head_qk_dim, self._head_v_dim = head_qk_dim, head_v_dim
        self._use_qk_l2norm = use_qk_l2norm
        qk_dim, v_dim = num_query_key_heads * head_qk_dim, num_value_heads * head_v_dim
        self._splits = [qk_dim, qk_dim, v_dim]
        self.qkv_proj = Linear(hidden_size, 2 * qk_dim + v_dim, bias=False)
        self.g_proj = Linear(hidden_size, v_dim, bias=False)
        self.b_proj = Linear(hidden_size, num_value_heads, bias=False)
        self.decay_gate = _build_decay_gate(decay_gate, hidden_size, num_value_heads)
        self.qkv_conv1d = CausalShortDepthwiseConv1d(2 * qk_dim + v_dim, conv_size)
        self.out_norm = RMSNorm(head_v_dim, eps=norm_eps)
        self.o_proj = Linear(v_dim, hidden_size, bias=False)

    def forward(self, hidden_states: torch.Tensor, attention_mask: torch.Tensor | None = None,
                position_embeddings: tuple[torch.Tensor, torch.Tensor] | None = None) -> torch.Tensor:
        del position_embeddings  # order is encoded by the recurrence itself; accepted so that decoder layers can treat
        # every token mixer alike
        b, s, _ = hidden_states.shape
        if attention_mask is not None:
            hidden_states = hidden_states * attention_mask.unsqueeze(-1).to(hidden_states.dtype)
        q, k, v = torch.split(self.qkv_conv1d(self.qkv_proj(hidden_states)), self._splits, dim=-1)
        decay = self.decay_gate(hidden_states)
        beta = torch.sigmoid(self.b_proj(hidden_states).float())
        q = q.reshape(b, s, self._num_qk_heads, self._head_qk_dim)
        k = k.reshape(b, s, self._num_qk_heads, self._head_qk_dim)
        v = v.reshape(b, s, self._num_v_heads, self._head_v_dim)
        if self._groups > 1:
            q = q.repeat_interleave(self._groups, dim=2)
            k = k.repeat_interleave(self._groups, dim=2)
        out = chunk_gated_delta_rule(q, k, v, decay, beta, use_qk_l2norm=self._use_qk_l2norm)
        out = self.out_norm(out.contiguous()).reshape(b, s, -1)
        return self.o_proj(silu_mul(self.g_proj(hidden_states), out))

    def reset_parameters(self) -> None:
        for m in (self.qkv_proj, self.g_proj, self.b_proj, self.decay_gate, self.o_proj, self.qkv_conv1d, self.out_norm):
            m.reset_parameters()
