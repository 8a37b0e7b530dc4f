from __future__ import annotations

import torch
from torch import nn

from d9d_b200.kernel.rope.fused_qk import fused_qk_norm_rope_supported, qk_norm_rope
from d9d_b200.module.base import ModuleLateInit
from d9d_b200.module.block.attention.sdpa import FlashSdpa
from d9d_b200.module.block.linear import Linear
from d9d_b200.module.block.normalization import RMSNorm
from d9d_b200.module.block.positional import RotaryEmbeddingApplicator, RotaryEmbeddingStyle


class GroupedQueryAttention(nn.Module, ModuleLateInit):
    """GQA: q/k/v projections -> optional per-head q/k RMSNorm -> (partial) RoPE -> SDPA -> optional sigmoid
    output gate -> o projection.  Parity: reference ``d9d/module/block/attention/grouped_query.py:10-171``.
    """

    def __init__(
        self,
        hidden_size: int,
        num_attention_heads: int,
        num_key_value_heads: int,
        head_dim: int,
        qk_norm_eps: float | None,
        is_causal: bool,
        rope_style: RotaryEmbeddingStyle,
        rope_dim: int | None = None,
        enable_output_gate: bool = False,
        qk_norm_zero_centered: bool = False,
    ) -> None:
        super().__init__()
        self._head_dim = head_dim
        self._num_key_value_groups = num_attention_heads // num_key_value_heads
        self._scaling = head_dim**-0.5
        self._rope_dim = rope_dim
        self._is_causal = is_causal

        q_out = num_attention_heads * head_dim
        kv_out = num_key_value_heads * head_dim
        self.q_proj = Linear(hidden_size, q_out, bias=False)
        self.gate_proj = Linear(hidden_size, q_out, bias=False) if enable_output_gate else None
        self.k_proj = Linear(hidden_size, kv_out, bias=False)
        self.v_proj = Linear(hidden_size, kv_out, bias=False)
        self.o_proj = Linear(q_out, hidden_size, bias=False)

        if qk_norm_eps is not None:
            self.q_norm: RMSNorm | None = RMSNorm(head_dim, eps=qk_norm_eps, zero_centered=qk_norm_zero_centered)
            self.k_norm: RMSNorm | None = RMSNorm(head_dim, eps=qk_norm_eps, zero_centered=qk_norm_zero_centered)
        else:
            self.q_norm = None
            self.k_norm = None

        self.rope = RotaryEmbeddingApplicator(style=rope_style)
        self.kernel = FlashSdpa()

    @property
    def head_dim(self) -> int:
        return self._head_dim

    def forward(self, hidden_states: torch.Tensor, attention_mask: torch.Tensor | None,
                position_embeddings: tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
        # shapes are taken from the projections: under tensor + sequence parallelism they return the tokens of the whole
        # tensor-parallel group for this rank's heads, while ``hidden_states`` holds only the local tokens
        q = self.q_proj(hidden_states)
        lead = q.shape[:-1]
        per_head = (*lead, -1, self._head_dim)
        q = q.view(per_head)
        k = self.k_proj(hidden_states).view(per_head)
        v = self.v_proj(hidden_states).view(per_head)
        cos, sin = position_embeddings
        if (self.q_norm is not None and self.k_norm is not None
                and fused_qk_norm_rope_supported(q, k, self.q_norm.weight, self.k_norm.weight, cos.shape[-1])):
            # per-head RMSNorm of q and k + rotary embedding in one kernel
            q, k = qk_norm_rope(q, k, self.q_norm.weight, self.k_norm.weight, cos, sin, self.q_norm.eps,
                                self.q_norm.zero_centered, self.rope.style_code)
        else:
            if self.q_norm is not None:
                q = self.q_norm(q)
            if self.k_norm is not None:
                k = self.k_norm(k)
            # the rope kernel rotates the first cos.shape[-1] dims of each head and passes the rest through,
            # so partial RoPE needs no split/cat
            q, k = self.rope(q, k, cos, sin)

        out = self.kernel(q, k, v, attention_mask=attention_mask, is_causal=self._is_causal, scale=self._scaling)
        out = out.reshape(*lead, -1)
        if self.gate_proj is not None:
            out = out * torch.sigmoid(self.gate_proj(hidden_states))
        return self.o_proj(out)

    def reset_parameters(self) -> None:
        for mod in (self.q_proj, self.k_proj, self.v_proj, self.gate_proj, self.o_proj, self.q_norm, self.k_norm):
            if mod is not None:
                mod.reset_parameters()
