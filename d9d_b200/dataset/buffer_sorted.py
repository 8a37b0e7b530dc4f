from __future__ import annotations

import pickle
import random
from typing import Any, Protocol, TypeVar

from torch.distributed.checkpoint.stateful import Stateful
from torch.utils.data import Dataset

_T_co = TypeVar("_T_co", covariant=True)


class DatasetImplementingSortKeyProtocol(Protocol[_T_co]):
    """Dataset that can tell a sort key (e.g. token count) of an item without materialising it."""

    def __len__(self) -> int: ...

    def sort_key(self, index: int) -> Any: ...

    def __getitem__(self, item: int) -> _T_co: ...


class BufferSortedDataset(Dataset[_T_co], Stateful):
    """Length bucketing with local randomness: windows of ``buffer_size`` items are sorted by key (random
    tie-break), cut into packs of ``pack_size``, the packs are shuffled and each pack is shuffled internally.
    Exactly resumable (RNG state + current window are checkpointed).

    Parity: reference ``d9d/dataset/buffer_sorted.py:40-133``.
    """

    def __init__(self, base_dataset: DatasetImplementingSortKeyProtocol[_T_co], buffer_size: int, pack_size: int,
                 init_seed: int | None = None):
        self._base = base_dataset
        self._buffer_size = buffer_size
        self._pack_size = pack_size
        self._rng = random.Random(init_seed ^ 0x105E7 if init_seed is not None else None)
        self._window: list[int] = []
        self._window_idx = -1

    def _load_window(self, window_idx: int) -> None:
        lo = window_idx * self._buffer_size
        hi = min(lo + self._buffer_size, len(self._base))
        keyed = sorted(((self._base.sort_key(i), self._rng.random(), i) for i in range(lo, hi)), key=lambda t: (t[0], t[1]))
        ordered = [i for _, _, i in keyed]
        packs = [ordered[p : p + self._pack_size] for p in range(0, len(ordered), self._pack_size)]
        self._rng.shuffle(packs)
        for pack in packs:
            self._rng.shuffle(pack)
        self._window = [i for pack in packs for i in pack]
        self._window_idx = window_idx

    def __getitem__(self, index: int) -> _T_co:
        window_idx, offset = divmod(index, self._buffer_size)
        if window_idx != self._window_idx:
            self._load_window(window_idx)
        return self._base[self._window[offset]]

    def __len__(self) -> int:
        return len(self._base)

    def state_dict(self) -> dict[str, Any]:
        state = {"seed": pickle.dumps(self._rng.getstate()), "buffer_idx": self._window_idx, "buffer_indices": list(self._window)}
        if isinstance(self._base, Stateful):
            state["base_dataset"] = self._base.state_dict()
        return state

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._rng.setstate(pickle.loads(state_dict["seed"]))  # noqa: S301
        self._window_idx = state_dict["buffer_idx"]
        self._window = list(state_dict["buffer_indices"])
        if isinstance(self._base, Stateful):
            self._base.load_state_dict(state_dict["base_dataset"])
