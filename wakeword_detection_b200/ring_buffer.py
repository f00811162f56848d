"""Host-side fixed-capacity ring with the interface of the reference's RingBuffer
(reference: spokestack/ring_buffer.py:9-130 == utils/tf_lite/ring_buffer.py).

Kept for interface compatibility of the drop-in classes (`sample_window`,
`frame_window`, `encode_window` are public attributes of WakewordTrigger); the device
path keeps its rings in HBM (csrc/filter.cu, StreamState).  State is (storage, write
index, item count) rather than two pointers, with the same observable behaviour:
capacity+1 slots, `rewind()` makes the last `capacity` slots readable, `seek(k)` drops
the k oldest, IndexError on overflow/underflow.
"""
from typing import Union

import numpy as np


class RingBuffer:
    def __init__(self, shape: list, dtype=np.float32) -> None:
        shape = list(shape)
        self._slots = int(shape[0]) + 1
        self._dtype = dtype
        self._store = np.empty([self._slots] + shape[1:], dtype=dtype)
        self._w = 0      # next slot to write
        self._n = 0      # readable items (the n slots before _w)

    @property
    def is_empty(self) -> bool:
        return self._n == 0

    @property
    def is_full(self) -> bool:
        return self._n == self._slots - 1

    @property
    def capacity(self) -> int:
        return self._slots - 1

    def rewind(self):
        self._n = self._slots - 1
        return self

    def reset(self):
        self._w = (self._w - self._n) % self._slots
        self._n = 0
        return self

    def fill(self, value: Union[int, float]):
        self._store.fill(value)
        self._n = self._slots - 1
        return self

    def seek(self, steps: int):
        self._n = (self._n - steps) % self._slots
        return self

    def write(self, item) -> None:
        if self.is_full:
            raise IndexError("Buffer is full")
        self._store[self._w] = item
        self._w = (self._w + 1) % self._slots
        self._n += 1

    def read(self) -> np.ndarray:
        if self.is_empty:
            raise IndexError("Buffer is empty")
        i = (self._w - self._n) % self._slots
        self._n -= 1
        return self._store[i:i + 1]

    def read_all(self) -> np.ndarray:
        self.rewind()
        start = (self._w - self._n) % self._slots
        idx = (start + np.arange(self._n)) % self._slots
        self._n = 0
        return self._store[idx].astype(self._dtype)
