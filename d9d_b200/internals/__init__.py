"""Runtime internals: gradient synchronisation / clipping, pipeline state, async metrics, profiling, seeding."""
