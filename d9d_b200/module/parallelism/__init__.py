"""Horizontal parallelism: composable ``parallelize_*`` hooks over DeviceMesh / DTensor."""
