"""Row-chunked logits helper shared by the forward / backward stand-ins."""

from __future__ import annotations

import math

import torch

CHUNK_BYTES = 1 << 30  # bf16 logits per chunk


def row_chunks(n_rows: int, vocab: int) -> int:
    return max(256, min(n_rows, CHUNK_BYTES // max(2 * vocab, 1) // 256 * 256)) if n_rows else 1


def chunk_logits(e_rows: torch.Tensor, c: torch.Tensor, bias: torch.Tensor | None, softcap: float | None):
    """Returns (fp32 logits after bias/softcap, tanh term for the softcap backward or None)."""
    logits = torch.mm(e_rows, c.t()).float()
    if bias is not None:
        logits += bias.float()
    t = None
    if softcap is not None and not math.isinf(softcap) and softcap != 0:
        t = torch.tanh(logits / softcap)
        logits = t * softcap
    return logits, t
