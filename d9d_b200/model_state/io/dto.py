from pydantic import BaseModel

MODEL_STATE_INDEX_FILE_NAME = "model.safetensors.index.json"


class ModelStateIndexMeta(BaseModel):
    total_size: int  # bytes of all tensors


class ModelStateIndex(BaseModel):
    """Contents of ``model.safetensors.index.json``: weight name -> shard file."""

    metadata: ModelStateIndexMeta
    weight_map: dict[str, str]
