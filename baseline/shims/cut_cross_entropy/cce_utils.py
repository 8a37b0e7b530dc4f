from __future__ import annotations

import enum
from dataclasses import asdict, dataclass


class LinearCrossEntropyImpl(enum.IntEnum):
    CCE = enum.auto()
    TORCH_COMPILE = enum.auto()
    CCE_EXACT = enum.auto()
    CCE_KAHAN = enum.auto()
    CCE_KAHAN_FULL_C = enum.auto()
    CCE_KAHAN_FULL_E = enum.auto()
    CCE_KAHAN_FULL = enum.auto()


@dataclass
class CCEPreset:
    filter_eps: float | str | None = "auto"
    accum_e_fp32: bool = False
    accum_c_fp32: bool = False
    filter_e_grad: bool = True
    filter_c_grad: bool = True


class CCEPresets:
    names = {"cce", "cce_exact", "cce_kahan", "cce_kahan_full_c", "cce_kahan_full_e", "cce_kahan_full"}

    @classmethod
    def build_for_impl(cls, impl: str, opts: CCEPreset) -> dict:
        if impl not in cls.names:
            raise ValueError(f"{impl!r} not in {cls.names}")
        out = asdict(opts)
        if impl == "cce_exact":
            out["filter_eps"] = None
        if impl.startswith("cce_kahan"):
            out["accum_e_fp32"] = out["accum_c_fp32"] = True
        if impl in ("cce_kahan_full_c", "cce_kahan_full"):
            out["filter_c_grad"] = False
        if impl in ("cce_kahan_full_e", "cce_kahan_full"):
            out["filter_e_grad"] = False
        return out
