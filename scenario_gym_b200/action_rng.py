"""
Device-side action source (``SgActionRng`` in include/sg_b200.h).

The "random accel/steer" configurations draw their ``VehicleAction`` rows from
``numpy.random.default_rng(seed)``.  ``ActionRng`` describes such a table by the generator state
and the position of the rows in the stream, so the CUDA kernels evaluate the same PCG64 stream in
place (bit-identical values, nothing to upload) while a host policy -- or the CPU oracle in the
tests -- keeps drawing the table with numpy itself.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple, Union

import numpy as np

from . import abi

_MASK64 = (1 << 64) - 1


@dataclass
class ActionRng:
    """
    Row k (k-th tick after reset), component c (0 accel, 1 steer), slot index i (= n*M + s) is
    ``low[c] + (high[c] - low[c]) * u_j`` with ``j = offset[c] + k * tick_stride + i`` and ``u_j`` the
    j-th double of ``Generator(PCG64).random()`` started from (``state``, ``inc``).
    """

    state: int
    inc: int
    offset: Tuple[int, int]
    tick_stride: int
    low: Tuple[float, float]
    high: Tuple[float, float]
    n_ticks: int
    nm: int

    @classmethod
    def from_generator(cls, rng: Union[int, np.random.Generator, np.random.PCG64], *, offset: Sequence[int],
                       tick_stride: int, low: Sequence[float], high: Sequence[float], n_ticks: int,
                       nm: int) -> "ActionRng":
        """From a seed, a ``Generator`` or a ``PCG64`` in the state the stream STARTS from (draw 0)."""
        if isinstance(rng, np.random.Generator):
            bg = rng.bit_generator
        elif isinstance(rng, np.random.PCG64):
            bg = rng
        else:
            bg = np.random.PCG64(rng)
        st = bg.state
        if st["bit_generator"] != "PCG64":
            raise TypeError("the device action source restates numpy's PCG64 only")
        return cls(state=int(st["state"]["state"]), inc=int(st["state"]["inc"]),
                   offset=(int(offset[0]), int(offset[1])), tick_stride=int(tick_stride),
                   low=(float(low[0]), float(low[1])), high=(float(high[0]), float(high[1])),
                   n_ticks=int(n_ticks), nm=int(nm))

    def struct(self) -> abi.SgActionRng:
        r = abi.SgActionRng()
        r.state_hi, r.state_lo = self.state >> 64, self.state & _MASK64
        r.inc_hi, r.inc_lo = self.inc >> 64, self.inc & _MASK64
        r.offset[0], r.offset[1] = self.offset
        r.tick_stride = self.tick_stride
        for c in range(2):
            r.low[c] = self.low[c]
            r.scale[c] = self.high[c] - self.low[c]  # numpy Generator.uniform: range = high - low
        return r

    def shard(self, lo_slot: int, n_slots: int) -> "ActionRng":
        """The same table restricted to slot indices [lo_slot, lo_slot + n_slots) (scenario shards)."""
        return ActionRng(self.state, self.inc, (self.offset[0] + lo_slot, self.offset[1] + lo_slot),
                         self.tick_stride, self.low, self.high, self.n_ticks, int(n_slots))

    def table(self, tick0: int = 0, n_ticks: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Rows [tick0, tick0 + n_ticks) as a (n_ticks, 2, nm) float64 table, drawn by numpy itself."""
        n_ticks = self.n_ticks - tick0 if n_ticks is None else n_ticks
        tab = np.empty((n_ticks, 2, self.nm)) if out is None else out
        for c in range(2):
            for k in range(n_ticks):
                bg = np.random.PCG64(0)
                bg.state = {"bit_generator": "PCG64", "state": {"state": self.state, "inc": self.inc},
                            "has_uint32": 0, "uinteger": 0}
                bg.advance(self.offset[c] + (tick0 + k) * self.tick_stride)
                tab[k, c] = np.random.Generator(bg).uniform(self.low[c], self.high[c], self.nm)
        return tab
