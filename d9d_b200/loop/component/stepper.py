from __future__ import annotations

from typing import Any

from torch.distributed.checkpoint.stateful import Stateful

from d9d_b200.loop.config import StepActionPeriod, StepActionSpecial


class Stepper(Stateful):
    """Current / total step counter with the ``should_do_action`` periodic-action rule (reference ``stepper.py``)."""

    def __init__(self, initial_step: int, total_steps: int):
        self._current = initial_step
        self._total = total_steps

    def step(self) -> None:
        self._current += 1

    @property
    def current_step(self) -> int:
        return self._current

    @property
    def total_steps(self) -> int:
        return self._total

    def state_dict(self) -> dict[str, Any]:
        return {"current_step": self._current, "total_steps": self._total}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        if state_dict["total_steps"] != self._total:
            raise ValueError(f"Step count differs: saved {state_dict['total_steps']}, current {self._total}. "
                             "Perhaps project configuration changed?")
        self._current = state_dict["current_step"]

    def should_do_action(self, action: StepActionPeriod, enable_on_last_step_if_periodic: bool = False,
                         is_post_step_action: bool = False) -> bool:
        """Evaluate a period for the step being executed (``current + 1``) or, for post-step actions, for the step
        counter as already advanced."""
        position = self._current + (0 if is_post_step_action else 1)
        if action == StepActionSpecial.disable:
            return False
        if action == StepActionSpecial.last_step:
            return position == self._total
        if isinstance(action, int) and not isinstance(action, bool):
            if action <= 0:
                raise ValueError("step period must be positive")
            return position % action == 0 or (enable_on_last_step_if_periodic and position == self._total)
        raise ValueError("Invalid step configuration")
