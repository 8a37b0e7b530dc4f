"""Model families (alias of :mod:`d9d_b200.module.model`)."""

from d9d_b200.module.model import *  # noqa: F401,F403
from d9d_b200.module.model import llama3, mixtral, qwen3_dense, qwen3_moe  # noqa: F401
