"""Context (sequence) parallel attention: the sequence dimension of a batch is split over a process group.

Net-new relative to the reference, which reserves the ``cp_shard`` mesh dimension but implements nothing on it
(``d9d/core/dist_context/device_mesh_domains.py:108-151``; models reject ``cp > 1``).  Two exchange patterns:

* :func:`ulysses_attention` - all-to-all that trades the sequence split for a head split, full-sequence attention on
  ``heads / world`` local heads with the regular (fast) attention kernels, all-to-all back.  On an NVSwitch node the
  all-to-all runs at full bisection bandwidth, so this is the default when the head counts allow it;
* :func:`ring_attention` - K/V blocks travel around a ring while every rank attends its local queries to the block it
  currently holds; partial results are merged with their log-sum-exps.  Works for any head count and keeps activation
  memory at ``O(S / world)``.

:mod:`.layout` defines which tokens a rank holds (contiguous or load-balanced *zig-zag* chunks).
"""

from .layout import ContextParallelLayout, gather_sequence, local_sequence_indices, shard_sequence
from .ring import ring_attention
from .ulysses import ulysses_attention, ulysses_supported

__all__ = [
    "ContextParallelLayout",
    "gather_sequence",
    "local_sequence_indices",
    "ring_attention",
    "shard_sequence",
    "ulysses_attention",
    "ulysses_supported",
]
