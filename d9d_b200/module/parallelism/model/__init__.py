"""Per-family parallelisation plans."""
