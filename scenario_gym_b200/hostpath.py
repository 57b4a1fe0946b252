"""
Host-buffer rollouts (``sg_rollout_host``): the call a user with inputs in HOST memory makes.

``HostRollout`` stages a packed scene (and, for table-fed vehicle scenes, the action table) in
pinned host memory once; every ``run()`` then copies the scene host->device, resets, rolls every
scenario out to ``is_done`` and copies the per-scenario results back -- the end-to-end path
``bench.py`` times.  With an ``ActionRng`` the actions are drawn inside the kernel and only the
scene crosses PCIe.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Union

import numpy as np
import torch

from . import abi
from .action_rng import ActionRng
from .engine import Engine

RESULT_FIELDS = (
    ("ego_avg_speed", torch.float64, ()), ("ego_max_speed", torch.float64, ()), ("ego_dist", torch.float64, ()),
    ("first_coll_tick", torch.int32, ()), ("first_coll_pair", torch.int32, (2,)), ("n_pair_ticks", torch.int64, ()),
    ("rss_flags", torch.uint8, ()), ("tick", torch.int32, ()), ("t", torch.float64, ()),
)


class HostRollout:
    """Pinned host mirrors of an engine's scene + action source, and the result buffers."""

    def __init__(self, engine: Engine, actions: Union[None, ActionRng, np.ndarray, torch.Tensor] = None,
                 device_union: bool = True):
        self.engine = eng = engine
        scene = eng.scene
        N = scene.N
        # The BatchReplayEntity union table is 6 M times the size of its knot times: with finite control
        # points (the reference's nan_to_num is then the identity) only the knot times are uploaded and
        # sg_rollout_host builds the rows on the device (sg_build_union_x), bit-identical to the host's.
        self.device_union = bool(device_union and scene.union_on_device_ok())
        self._host = {k: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                      for k, a in scene.arrays(union_rows=not self.device_union).items()}
        self._hs = abi.SgScene()
        for f, _ in abi.SgScene._fields_:
            setattr(self._hs, f, getattr(eng._sc, f))
        for k, t in self._host.items():
            setattr(self._hs, k, t.data_ptr() if t.numel() else None)
        if self.device_union:
            self._hs.union_x = None
        self._hin, self._din = abi.SgInputs(), abi.SgInputs()
        self.action_host: Optional[torch.Tensor] = None
        if isinstance(actions, ActionRng):
            if actions.nm != N * scene.M:
                raise ValueError(f"ActionRng describes {actions.nm} slots, the scene has {N * scene.M}")
            self._hin.use_rng, self._hin.rng_tick0, self._hin.n_action_ticks = 1, 0, actions.n_ticks
            self._hin.rng = actions.struct()
        elif actions is not None:
            a = torch.as_tensor(actions)
            if a.dtype != torch.float32:
                a = a.to(torch.float64)
            if a.dim() != 3 or tuple(a.shape[1:]) != (2, N * scene.M):
                raise ValueError(f"actions must be (T, 2, {N * scene.M})")
            self.action_host = a.cpu().contiguous().pin_memory() if not (a.device.type == "cpu" and a.is_pinned()) else a
            self._action_dev = torch.empty(tuple(a.shape), dtype=a.dtype, device=eng.device)
            if a.dtype == torch.float32:
                self._hin.actions_f32, self._din.actions_f32 = self.action_host.data_ptr(), self._action_dev.data_ptr()
            else:
                self._hin.actions, self._din.actions = self.action_host.data_ptr(), self._action_dev.data_ptr()
            self._hin.n_action_ticks = self._din.n_action_ticks = a.shape[0]
        self.results: Dict[str, torch.Tensor] = {
            name: torch.empty((N,) + shape, dtype=dt, pin_memory=True) for name, dt, shape in RESULT_FIELDS}
        self.results["event_count"] = torch.empty(1, dtype=torch.int32, pin_memory=True)
        self._res = abi.SgHostResults()
        for k, t in self.results.items():
            setattr(self._res, k, t.data_ptr())
        lib = eng.lib
        self.h2d_bytes = int(lib["host_h2d_bytes"](C.byref(self._hs), C.byref(self._hin), 1))
        self.d2h_bytes = int(lib["host_d2h_bytes"](C.byref(self._hs)))

    def launch(self, copy_static: bool = True) -> None:
        """Enqueue scene H2D (+ chunked action H2D), reset, rollout and result D2H on the current stream."""
        eng = self.engine
        eng._check(eng.lib["rollout_host"](
            C.byref(self._hs), C.byref(eng._sc), C.byref(eng.params), C.byref(eng._st), C.byref(self._hin),
            C.byref(self._din), C.byref(self._res), int(copy_static), eng.dev_index, eng._stream()))

    def run(self, copy_static: bool = True) -> Dict[str, np.ndarray]:
        """``launch()`` and wait; returns numpy views of the pinned result buffers."""
        self.launch(copy_static)
        torch.cuda.current_stream(self.engine.device).synchronize()
        return {k: t.numpy() for k, t in self.results.items()}
