"""Program construction passes (reference ``d9d/pipelining/infra/schedule/component/program/__init__.py``)."""

from d9d_b200.pipelining.infra.communications import add_communication_ops, validate_program
from d9d_b200.pipelining.infra.programs import PipelineProgramBuilder, program_to_text
from d9d_b200.pipelining.infra.simulator import SchedulePolicy, simulate
from d9d_b200.pipelining.infra.topology import (
    ScheduleStyle,
    build_stage_to_host_rank_topology,
    invert_stage_to_host_rank_topology,
)

__all__ = [
    "PipelineProgramBuilder",
    "SchedulePolicy",
    "ScheduleStyle",
    "add_communication_ops",
    "build_stage_to_host_rank_topology",
    "invert_stage_to_host_rank_topology",
    "program_to_text",
    "simulate",
    "validate_program",
]
