"""Core layer: distributed context / mesh domains, collectives helpers, pytree sharding, autograd direction switch."""
