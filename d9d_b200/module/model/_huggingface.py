"""HuggingFace ⇄ native checkpoint key correspondences, described once and compiled in either direction.

The reference ships four hand-written mapper factories per direction per family
(``d9d/module/model/qwen3_moe/huggingface.py:28-412``, dense twin under ``qwen3_dense/``).  Here every family is a
short list of *rules* (data); :func:`compile_rules` turns a rule list into a ``ModelStateMapper`` DAG for the requested
:class:`Direction`.  Native names are the left column of the reference's format so checkpoints interoperate.
"""

from __future__ import annotations

import dataclasses
import enum
from collections.abc import Callable, Sequence

import torch

from d9d_b200.model_state.mapper import ModelStateMapper, StateGroup
from d9d_b200.model_state.mapper.compose import (
    ModelStateMapperParallel,
    ModelStateMapperPrefixScope,
    ModelStateMapperSequential,
)
from d9d_b200.model_state.mapper.leaf import (
    ModelStateMapperChunkTensors,
    ModelStateMapperConcatenateTensors,
    ModelStateMapperIdentity,
    ModelStateMapperRename,
    ModelStateMapperSqueeze,
    ModelStateMapperStackTensors,
    ModelStateMapperTranspose,
    ModelStateMapperUnsqueeze,
    ModelStateMapperUnstackTensors,
)


class Direction(enum.StrEnum):
    FROM_HF = "from_huggingface"
    TO_HF = "to_huggingface"


class ExpertsFormat(enum.StrEnum):
    """How HuggingFace stores MoE experts: one ``nn.Linear`` per expert and projection (transformers 4.x) or fused 3-D
    tensors ``gate_up_proj [E, 2I, H]`` / ``down_proj [E, H, I]`` (transformers 5.x)."""

    MODULE_LIST = "module_list"
    FUSED = "fused"


@dataclasses.dataclass(frozen=True)
class Same:
    name: str


@dataclasses.dataclass(frozen=True)
class Renamed:
    hf: str
    native: str


@dataclasses.dataclass(frozen=True)
class ExpertsPerModule:
    """HF: ``hf_pattern.format(e=expert, proj=hf_proj)`` are ``[out, in]`` matrices; native: ``[E, in, out]``."""

    hf_pattern: str
    projections: tuple[tuple[str, str], ...]  # (hf projection name, native projection name)
    native_pattern: str
    num_experts: int


@dataclasses.dataclass(frozen=True)
class ExpertsFused:
    hf_gate_up: str
    hf_down: str
    native_gate: str
    native_up: str
    native_down: str


@dataclasses.dataclass(frozen=True)
class Scoped:
    hf_prefix: str
    native_prefix: str
    rules: tuple["Rule", ...]


@dataclasses.dataclass(frozen=True)
class HeadInterleaved:
    """HF fuses several per-head projections into one weight ``[heads * len(natives) * head_dim, in]`` laid out head by
    head (``[part_0 | part_1 | ...]`` inside every head, e.g. Qwen3.5's query + output gate); native: one weight per part."""

    hf: str
    natives: tuple[str, ...]
    head_dim: int


@dataclasses.dataclass(frozen=True)
class Unsqueezed:
    """HF keeps an extra singleton dim (``conv1d.weight [C, 1, K]``), native does not (``[C, K]``)."""

    hf: str
    native: str
    dim: int


Rule = Same | Renamed | ExpertsPerModule | ExpertsFused | Scoped | HeadInterleaved | Unsqueezed


class _HeadInterleaveMapper(ModelStateMapper):
    def __init__(self, rule: HeadInterleaved, from_hf: bool):
        self._rule, self._from_hf = rule, from_hf

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        fused, parts = frozenset([self._rule.hf]), frozenset(self._rule.natives)
        return frozenset([StateGroup(inputs=fused, outputs=parts) if self._from_hf else StateGroup(inputs=parts, outputs=fused)])

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        n, d = len(self._rule.natives), self._rule.head_dim
        if self._from_hf:
            fused = group[self._rule.hf]
            per_head = fused.reshape(-1, n, d, fused.shape[-1])  # [heads, part, head_dim, in]
            return {name: per_head[:, i].reshape(-1, fused.shape[-1]).contiguous() for i, name in enumerate(self._rule.natives)}
        parts = [group[name] for name in self._rule.natives]
        stacked = torch.stack([p.reshape(-1, d, p.shape[-1]) for p in parts], dim=1)  # [heads, part, head_dim, in]
        return {self._rule.hf: stacked.reshape(-1, parts[0].shape[-1]).contiguous()}


def _compile_one(rule: Rule, direction: Direction) -> list[ModelStateMapper]:
    from_hf = direction == Direction.FROM_HF
    if isinstance(rule, Same):
        return [ModelStateMapperIdentity(rule.name)]
    if isinstance(rule, Renamed):
        return [ModelStateMapperRename(rule.hf, rule.native) if from_hf else ModelStateMapperRename(rule.native, rule.hf)]
    if isinstance(rule, ExpertsPerModule):
        out = []
        for hf_proj, native_proj in rule.projections:
            hf_names = [rule.hf_pattern.format(e=e, proj=hf_proj) for e in range(rule.num_experts)]
            native = rule.native_pattern.format(proj=native_proj)
            if from_hf:
                out.append(ModelStateMapperSequential([
                    ModelStateMapperStackTensors(source_names=hf_names, target_name=native, dim=0),
                    ModelStateMapperTranspose(native, dims=(-1, -2)),
                ]))
            else:
                out.append(ModelStateMapperSequential([
                    ModelStateMapperTranspose(native, dims=(-1, -2)),
                    ModelStateMapperUnstackTensors(source_name=native, target_names=hf_names, dim=0),
                ]))
        return out
    if isinstance(rule, ExpertsFused):
        if from_hf:
            return [
                ModelStateMapperSequential([
                    ModelStateMapperTranspose(rule.hf_gate_up, dims=(-1, -2)),
                    ModelStateMapperChunkTensors(source_name=rule.hf_gate_up, target_names=[rule.native_gate, rule.native_up], dim=-1),
                ]),
                ModelStateMapperSequential([
                    ModelStateMapperTranspose(rule.hf_down, dims=(-1, -2)),
                    ModelStateMapperRename(rule.hf_down, rule.native_down),
                ]),
            ]
        return [
            ModelStateMapperSequential([
                ModelStateMapperConcatenateTensors(source_names=[rule.native_gate, rule.native_up], target_name=rule.hf_gate_up, dim=-1),
                ModelStateMapperTranspose(rule.hf_gate_up, dims=(-1, -2)),
            ]),
            ModelStateMapperSequential([
                ModelStateMapperRename(rule.native_down, rule.hf_down),
                ModelStateMapperTranspose(rule.hf_down, dims=(-1, -2)),
            ]),
        ]
    if isinstance(rule, HeadInterleaved):
        return [_HeadInterleaveMapper(rule, from_hf)]
    if isinstance(rule, Unsqueezed):
        if from_hf:
            return [ModelStateMapperSequential([ModelStateMapperSqueeze(rule.hf, rule.dim), ModelStateMapperRename(rule.hf, rule.native)])]
        return [ModelStateMapperSequential([ModelStateMapperRename(rule.native, rule.hf), ModelStateMapperUnsqueeze(rule.hf, rule.dim)])]
    if isinstance(rule, Scoped):
        inner = compile_rules(rule.rules, direction)
        src, dst = (rule.hf_prefix, rule.native_prefix) if from_hf else (rule.native_prefix, rule.hf_prefix)
        return [ModelStateMapperPrefixScope(inner, source_prefix=src, target_prefix=dst)]
    raise TypeError(f"unknown rule {rule!r}")


def compile_rules(rules: Sequence[Rule], direction: Direction) -> ModelStateMapper:
    mappers: list[ModelStateMapper] = []
    for rule in rules:
        mappers.extend(_compile_one(rule, direction))
    return ModelStateMapperParallel(mappers)


# ------------------------------------------------------------------------------------------ shared rule builders
def single_vocab_name(split_vocab_order: list[str]) -> str:
    if len(split_vocab_order) != 1:
        raise ValueError("HuggingFace mappers can only process a single vocab split")
    return split_vocab_order[0]


def attention_rules(qk_norm: bool) -> tuple[Rule, ...]:
    names = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj"]
    if qk_norm:
        names += ["self_attn.q_norm", "self_attn.k_norm"]
    return tuple(Same(f"{n}.weight") for n in names)


def norm_rules() -> tuple[Rule, ...]:
    return (Same("input_layernorm.weight"), Same("post_attention_layernorm.weight"))


def dense_mlp_rules() -> tuple[Rule, ...]:
    return tuple(Same(f"mlp.{p}.weight") for p in ("gate_proj", "up_proj", "down_proj"))


def backbone_rules(layer_rules: "tuple[Rule, ...] | Callable[[int], tuple[Rule, ...]]", num_layers: int,
                   vocab_name: str) -> tuple[Rule, ...]:
    """``layer_rules``: the rules of one decoder layer, or a function of the layer index for heterogeneous stacks."""
    per_layer = layer_rules if callable(layer_rules) else (lambda _index: layer_rules)
    return (
        Renamed("embed_tokens.weight", f"embed_tokens.token_embedding.{vocab_name}.weight"),
        *(Scoped(f"layers.{i}.", f"layers.{i}.", per_layer(i)) for i in range(num_layers)),
        Same("norm.weight"),
    )


def latent_attention_rules(low_rank_query: bool) -> tuple[Rule, ...]:
    """DeepSeek multi-head latent attention: HF ``kv_a_proj_with_mqa / kv_a_layernorm / kv_b_proj`` (and the optional
    ``q_a_proj / q_a_layernorm / q_b_proj`` bottleneck) <-> ``kv_down_proj / kv_down_norm / kv_up_proj`` (``q_proj.*``)."""
    query = ((Renamed("self_attn.q_a_proj.weight", "self_attn.q_proj.down_proj.weight"),
              Renamed("self_attn.q_a_layernorm.weight", "self_attn.q_proj.norm.weight"),
              Renamed("self_attn.q_b_proj.weight", "self_attn.q_proj.up_proj.weight"))
             if low_rank_query else (Same("self_attn.q_proj.weight"),))
    return (*query,
            Renamed("self_attn.kv_a_proj_with_mqa.weight", "self_attn.kv_down_proj.weight"),
            Renamed("self_attn.kv_a_layernorm.weight", "self_attn.kv_down_norm.weight"),
            Renamed("self_attn.kv_b_proj.weight", "self_attn.kv_up_proj.weight"),
            Same("self_attn.o_proj.weight"))


def shared_expert_rules(hf_name: str = "mlp.shared_experts") -> tuple[Rule, ...]:
    return tuple(Renamed(f"{hf_name}.{p}.weight", f"mlp.shared_expert.expert.{p}.weight") for p in ("gate_proj", "up_proj", "down_proj"))


def causal_lm_rules(backbone: tuple[Rule, ...], vocab_name: str) -> tuple[Rule, ...]:
    return (Scoped("model.", "model.", backbone), Renamed("lm_head.weight", f"lm_head.lm_head.{vocab_name}.weight"))


def classification_rules(backbone: tuple[Rule, ...]) -> tuple[Rule, ...]:
    return (Scoped("model.", "model.", backbone), Renamed("score.weight", "cls_head.score.weight"))


def embedding_rules(backbone: tuple[Rule, ...]) -> tuple[Rule, ...]:
    # HF embedding checkpoints are bare backbones (no ``model.`` prefix)
    return (Scoped("", "model.", backbone),)
