"""Pipeline runtime: action IR, schedule programs, stage runtime, executors."""
