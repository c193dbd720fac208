"""Named views over flat solution arrays (reference: Solverz/variable/variables.py:69-296,
utilities/address.py).  A batched solution has a leading scenario axis: ``array[B, n]``."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

__all__ = ["Address", "Vars", "TimeVars"]


class Address:
    def __init__(self):
        self._slices: Dict[str, slice] = {}
        self.total_size = 0

    def add(self, name: str, size: int):
        if name in self._slices:
            raise ValueError(f"Variable {name} already exists!")
        self._slices[name] = slice(self.total_size, self.total_size + int(size))
        self.total_size += int(size)

    @property
    def names(self) -> List[str]:
        return list(self._slices)

    @property
    def v(self) -> Dict[str, np.ndarray]:
        return {k: np.arange(s.start, s.stop) for k, s in self._slices.items()}

    def __getitem__(self, name: str) -> slice:
        return self._slices[name]

    def __contains__(self, name):
        return name in self._slices

    def __eq__(self, other):
        return isinstance(other, Address) and self._slices == other._slices


class Vars:
    """Flat vector (or ``[B, n]`` batch) with named slices."""

    def __init__(self, a: Address, array):
        self.a = a
        array = np.asarray(array, dtype=float)
        if array.shape[-1] != a.total_size:
            raise ValueError("Unequal address size and array size!")
        self.array = array

    @property
    def var_list(self):
        return self.a.names

    @property
    def var_size(self):
        return {k: self.a[k].stop - self.a[k].start for k in self.a.names}

    @property
    def total_size(self):
        return self.a.total_size

    def __getitem__(self, item):
        if isinstance(item, str):
            return self.array[..., self.a[item]]
        return self.array[item]

    def __setitem__(self, key, value):
        if key not in self.a:
            raise ValueError(f"There is no variable {key}!")
        self.array[..., self.a[key]] = value

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.array, dtype=dtype)

    def __repr__(self):
        return f"Variables (size {self.total_size}) {self.var_list}"


class TimeVars:
    """Trajectory ``[nt, n]`` (or ``[B, nt, n]``) with named slices (variables.py:196-296)."""

    def __init__(self, a: Address, array):
        self.a = a
        self.array = np.asarray(array, dtype=float)

    def __getitem__(self, item):
        if isinstance(item, str):
            return self.array[..., self.a[item]]
        if self.array.ndim == 2:                       # one trajectory: variables.py:228-246
            if isinstance(item, (int, np.integer)):
                return Vars(self.a, self.array[item])
            if isinstance(item, slice):
                return TimeVars(self.a, self.array[item])
        return self.array[item]

    @property
    def var_list(self):
        return self.a.names

    @property
    def var_size(self):
        return {k: self.a[k].stop - self.a[k].start for k in self.a.names}

    def __len__(self):
        return self.array.shape[-2]

    @property
    def len(self):
        return self.array.shape[-2]

    @property
    def T(self):
        return self.array.T

    def append(self, other: "TimeVars"):
        if self.a != other.a:
            raise ValueError("Cannot concatenate two TimeVars with different variable addresses!")
        self.array = np.concatenate([self.array, other.array], axis=-2)
