from torch import nn

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel

from .base import PeftMethod


def inject_peft_and_freeze(method: PeftMethod, module: nn.Module) -> ModelStateMapper:
    """Freeze everything, inject the method, unfreeze what it returns.  The returned mapper redirects the keys of a
    stock checkpoint to the modified structure (e.g. ``x.weight -> x.base.weight``)."""
    for p in module.parameters():
        p.requires_grad = False
    result = method.inject(module)
    for p in result.parameters_to_train:
        p.requires_grad = True
    return ModelStateMapperParallel(result.load_state_mappers)


def merge_peft(method: PeftMethod, module: nn.Module) -> None:
    """Fold trained adapters back into the base weights and restore the original module structure."""
    method.merge(module)
