"""Stage -> pipeline-rank placement (reference ``component/program/topology.py:18-74``)."""

from __future__ import annotations

import enum


class ScheduleStyle(enum.StrEnum):
    loop = "loop"  # stage s lives on rank s mod P
    v = "v"  # zig-zag: 0..P-1 then P-1..0 (the first and the last stage share rank 0)


def build_stage_to_host_rank_topology(num_stages: int, pp_size: int, style: ScheduleStyle) -> dict[int, int]:
    if num_stages % pp_size != 0:
        raise ValueError(f"num_stages ({num_stages}) must be a multiple of the pipeline size ({pp_size})")
    if style == ScheduleStyle.loop:
        return {s: s % pp_size for s in range(num_stages)}
    if style == ScheduleStyle.v:
        out = {}
        for s in range(num_stages):
            lap, pos = divmod(s, pp_size)
            out[s] = pos if lap % 2 == 0 else pp_size - 1 - pos
        return out
    raise ValueError(f"unknown topology style {style}")


def invert_stage_to_host_rank_topology(stage_to_host: dict[int, int]) -> dict[int, list[int]]:
    out: dict[int, list[int]] = {}
    for stage in sorted(stage_to_host):
        out.setdefault(stage_to_host[stage], []).append(stage)
    return out
