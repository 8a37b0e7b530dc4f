from typing import Literal

from pydantic import BaseModel


class AimConfig(BaseModel):
    """Aim backend settings (reference ``d9d/tracker/provider/aim/config.py``)."""

    provider: Literal["aim"] = "aim"
    repo: str
    log_system_params: bool = True
    capture_terminal_logs: bool = True
    system_tracking_interval: int = 10
