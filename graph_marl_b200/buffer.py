"""Auto-growing log buffer, mirrors src/buffer.py:4-63 (host-side logging only)."""
import numpy as np


class Buffer:
    def __init__(self, capacity: int, shape, dtype) -> None:
        self._data = np.empty((capacity, *shape), dtype=dtype)
        self._capacity = capacity
        self._count = 0

    def insert(self, a):
        if isinstance(a, list):
            for e in a:
                self._insert_element(e)
        else:
            self._insert_element(a)

    def _insert_element(self, elem):
        if self._count == self._capacity:
            self._data = np.concatenate((self._data, self._data), axis=0)
            self._capacity = self._data.shape[0]
        self._data[self._count] = elem
        self._count += 1

    def get(self) -> np.ndarray:
        return self._data[: self._count]

    def clear(self):
        self._count = 0

    def mean(self, default=0):
        return self.get().mean() if self._count > 0 else default
