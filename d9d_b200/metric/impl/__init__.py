"""Concrete metrics."""
