"""Growable host-side log buffer with the interface of the reference's `Buffer`
(src/buffer.py:4-63, used at src/main.py:615-618, 742-747, 783-788): `insert`, `get`, `clear`,
`mean` and the `_count` attribute that main.py reads.

Own implementation: values are appended in blocks (a python list of values is written with one
vectorised assignment), storage grows to the next power of two that fits, and CUDA tensors
(batched rewards / device-side counters of the B200 path) are accepted and brought to the host
once per insert.
"""
import numpy as np


def _to_host(value):
    if hasattr(value, "detach") and hasattr(value, "cpu"):  # torch tensor, possibly on the GPU
        return value.detach().cpu().numpy()
    return value


class Buffer:
    def __init__(self, capacity: int, shape, dtype) -> None:
        self._shape = tuple(shape)
        self._store = np.zeros((max(int(capacity), 1),) + self._shape, dtype=dtype)
        self._count = 0

    def _reserve(self, extra: int) -> None:
        need = self._count + extra
        have = self._store.shape[0]
        if need <= have:
            return
        grown = np.zeros((1 << (need - 1).bit_length(),) + self._shape, dtype=self._store.dtype)
        grown[: self._count] = self._store[: self._count]
        self._store = grown

    def insert(self, a) -> None:
        """Append one element, or every element of a python list (main.py passes info lists)."""
        if isinstance(a, list):
            if not a:
                return
            block = np.asarray([_to_host(e) for e in a], dtype=self._store.dtype).reshape((len(a),) + self._shape)
        else:
            block = np.asarray(_to_host(a), dtype=self._store.dtype)
            block = np.broadcast_to(block, self._shape).reshape((1,) + self._shape)
        self._reserve(block.shape[0])
        self._store[self._count:self._count + block.shape[0]] = block
        self._count += block.shape[0]

    def get(self) -> np.ndarray:
        return self._store[: self._count]

    def clear(self) -> None:
        self._count = 0

    def mean(self, default=0):
        if self._count == 0:
            return default
        return self._store[: self._count].mean()
