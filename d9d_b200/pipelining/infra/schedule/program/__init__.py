"""Pipeline schedule implementations (reference ``d9d/pipelining/infra/schedule/program/__init__.py``)."""

from d9d_b200.pipelining.infra.programs import (
    DualPipeVPipelineProgramBuilder,
    Interleaved1F1BPipelineProgramBuilder,
    LoopedBFSPipelineProgramBuilder,
    ZeroBubbleVPipelineProgramBuilder,
)

__all__ = [
    "DualPipeVPipelineProgramBuilder",
    "Interleaved1F1BPipelineProgramBuilder",
    "LoopedBFSPipelineProgramBuilder",
    "ZeroBubbleVPipelineProgramBuilder",
]
