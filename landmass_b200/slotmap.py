"""`slotmap::DenseSlotMap` semantics (iteration order = dense storage order with
swap-remove), which define agent order, island scan order and the order of
`pathing_results` in the reference (slotmap 1.0.7)."""
from __future__ import annotations


class Key(tuple):
    """(idx, version) — ordered by idx then version like slotmap::KeyData."""

    __slots__ = ()

    def __new__(cls, idx: int, version: int):
        return super().__new__(cls, (idx, version))

    @property
    def idx(self) -> int:
        return self[0]

    @property
    def version(self) -> int:
        return self[1]

    def __repr__(self):
        return f"Key({self[0]}v{self[1]})"


class DenseSlotMap:
    def __init__(self):
        # slot 0 is slotmap's sentinel
        self._slots = [[0, 0]]  # [version, dense index or next free]
        self._free_head = 1
        self._keys: list[Key] = []
        self._values: list = []

    def insert(self, value) -> Key:
        dense = len(self._keys)
        idx = self._free_head
        if idx < len(self._slots):
            slot = self._slots[idx]
            self._free_head = slot[1]
            slot[0] |= 1  # occupied versions are odd
            slot[1] = dense
        else:
            self._slots.append([1, dense])
            self._free_head = len(self._slots)
        key = Key(idx, self._slots[idx][0])
        self._keys.append(key)
        self._values.append(value)
        return key

    def _dense(self, key: Key):
        if key.idx >= len(self._slots):
            return None
        slot = self._slots[key.idx]
        if slot[0] != key.version:
            return None
        return slot[1]

    def remove(self, key: Key):
        dense = self._dense(key)
        if dense is None:
            return None
        slot = self._slots[key.idx]
        slot[0] += 1  # even = vacant
        slot[1] = self._free_head
        self._free_head = key.idx
        value = self._values[dense]
        last_key = self._keys[-1]
        self._keys[dense] = last_key
        self._values[dense] = self._values[-1]
        self._keys.pop()
        self._values.pop()
        if dense < len(self._keys):
            self._slots[last_key.idx][1] = dense
        return value

    def get(self, key: Key):
        dense = self._dense(key)
        return None if dense is None else self._values[dense]

    def __contains__(self, key) -> bool:
        return self._dense(key) is not None

    def keys(self):
        return list(self._keys)

    def values(self):
        return list(self._values)

    def items(self):
        return list(zip(self._keys, self._values))

    def __len__(self):
        return len(self._keys)

    def dense_index(self, key: Key):
        return self._dense(key)
