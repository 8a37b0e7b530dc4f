"""Pipeline parallelism: schedule programs, stage runtime, executors."""
