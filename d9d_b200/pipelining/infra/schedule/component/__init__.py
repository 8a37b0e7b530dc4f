"""Building blocks of schedules: program construction passes and the runtime that executes programs."""
