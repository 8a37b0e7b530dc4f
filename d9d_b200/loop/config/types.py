import enum


class StepActionSpecial(enum.StrEnum):
    last_step = "last_step"  # exactly once, at the very end of the run
    disable = "disable"  # never


StepActionPeriod = int | StepActionSpecial
"""Either a period in steps or a special flag."""
