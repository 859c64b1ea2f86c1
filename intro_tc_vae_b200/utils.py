"""Writer / loss-dict helpers the TC solvers touch (reference ``utils.py:48-74``)."""
from __future__ import annotations

from typing import Union


class LossDict(dict):
    """Dict of running losses: ``a + b`` adds key-wise (missing keys count as 0), ``a / k`` scales
    (reference utils.py:48-60; used for the last-epoch averages at train.py:194,222,246)."""

    def __add__(self, other: "LossDict") -> "LossDict":
        out = LossDict()
        for key in sorted(set(self) | set(other)):
            out[key] = self.get(key, 0) + other.get(key, 0)
        return out

    def __truediv__(self, value: Union[int, float]) -> "LossDict":
        return LossDict((k, v / value) for k, v in self.items())


class SingletonWriter:
    """Process-wide holder of the TensorBoard writer and the current iteration (reference utils.py:62-74).
    ``writer`` / ``cur_iter`` / ``test_iter`` are plain attributes set by the training driver."""

    writer = None
    cur_iter = 0
    test_iter = 1

    def __new__(cls):
        if not hasattr(cls, "instance"):
            cls.instance = super().__new__(cls)
        return cls.instance

    @property
    def write_test_iter(self):
        return self.writer and self.cur_iter % self.test_iter == 0
