"""Model building blocks, model families and horizontal-parallelism plans."""
