"""Build simple mappers from modules / other mappers (reference ``d9d/model_state/mapper/adapters``)."""

from torch import nn

from d9d_b200.model_state.mapper.abc import ModelStateMapper
from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel
from d9d_b200.model_state.mapper.leaf import ModelStateMapperIdentity


def identity_mapper_from_module(module: nn.Module) -> ModelStateMapper:
    """Pass-through for every key of ``module.state_dict()`` (plain ``load_state_dict`` behaviour)."""
    return ModelStateMapperParallel([ModelStateMapperIdentity(key) for key in module.state_dict()])


def identity_mapper_from_mapper_outputs(mapper: ModelStateMapper) -> ModelStateMapper:
    """Pass-through for every key the given mapper produces."""
    return ModelStateMapperParallel([ModelStateMapperIdentity(key) for key in sorted(mapper.all_outputs())])


__all__ = ["identity_mapper_from_mapper_outputs", "identity_mapper_from_module"]
