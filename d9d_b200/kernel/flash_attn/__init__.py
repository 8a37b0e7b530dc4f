from .function import attention_reference, flash_attn_func, flash_attn_varlen_func

__all__ = ["attention_reference", "flash_attn_func", "flash_attn_varlen_func"]
