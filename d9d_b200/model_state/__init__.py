"""Model-state IO: declarative mapper DAGs + streaming safetensors reader/writer (reference ``d9d/model_state``)."""
