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


def restrict_mapper_to_module(mapper: ModelStateMapper, module: nn.Module, *, module_keys_are: str) -> ModelStateMapper:
    """The part of a whole-model mapper that touches ``module``'s own state - what a pipeline stage needs.

    ``module_keys_are="outputs"``: an import mapper (checkpoint keys -> module keys) keeps the groups whose outputs all
    belong to the module; ``"inputs"``: an export mapper (module keys -> checkpoint keys) keeps the groups whose inputs all
    belong to it.  A group that straddles the stage boundary is an error.
    """
    from d9d_b200.model_state.mapper.compose import ModelStateMapperSelectGroups

    if module_keys_are not in ("inputs", "outputs"):
        raise ValueError("module_keys_are must be 'inputs' or 'outputs'")
    own = set(module.state_dict())

    def belongs(group) -> bool:  # noqa: ANN001
        keys = getattr(group, module_keys_are)
        if keys <= own:
            return True
        if keys & own:
            raise ValueError(f"mapper group {sorted(keys)} is split between this module and another pipeline stage")
        return False

    return ModelStateMapperSelectGroups(mapper, belongs)


__all__ = ["identity_mapper_from_mapper_outputs", "identity_mapper_from_module", "restrict_mapper_to_module"]
