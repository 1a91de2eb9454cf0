"""Replay tapes on the device (the `adc_tape` of include/adcraft_b200.h).

A tape holds, for every (env, keyword) unit of ONE env step, the pre-drawn values in the order
the reference consumes them: the volume, one competitor bid per auction, one
uniform per click slot, one per accepted click, one revenue per conversion; for explicit
keywords the per-sub-step impression counts and one cost per impression; optional drift
coefficients.  Streams are CSR over the E*K units.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _capi


@dataclass
class DeviceTape:
    volume: torch.Tensor
    comp_off: Optional[torch.Tensor]
    comp_cents: Optional[torch.Tensor]
    click_off: torch.Tensor
    u_click: torch.Tensor
    conv_off: torch.Tensor
    u_conv: torch.Tensor
    rev_off: torch.Tensor
    rev_cents: torch.Tensor
    impr: Optional[torch.Tensor] = None
    cost_off: Optional[torch.Tensor] = None
    cost: Optional[torch.Tensor] = None
    drift: Optional[torch.Tensor] = None

    def c_struct(self) -> _capi.Tape:
        t = _capi.Tape()
        for name, _ in _capi.Tape._fields_:
            v = getattr(self, name)
            setattr(t, name, None if v is None else v.data_ptr())
        return t

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.__dict__.values() if v is not None)

    @staticmethod
    def from_host(env_tapes: Sequence[object], device) -> "DeviceTape":
        """Concatenate per-env host tapes (objects with numpy attributes volume, comp_off,
        comp_cents, click_off, u_click, conv_off, u_conv, rev_off, rev_cents and optionally impr,
        cost_off, cost, drift -- per-env CSR over its K keywords) into one batch tape."""
        def cat_csr(off_name, val_name, dtype):
            offs, vals, base = [np.zeros(1, np.int64)], [], 0
            for t in env_tapes:
                off = np.asarray(getattr(t, off_name), np.int64)
                val = np.asarray(getattr(t, val_name), dtype)
                offs.append(off[1:] + base)
                vals.append(val[: off[-1]])
                base += int(off[-1])
            flat = np.concatenate(vals) if vals else np.zeros(0, dtype)
            if flat.size == 0:
                flat = np.zeros(1, dtype)  # keep a valid device pointer
            return (torch.from_numpy(np.concatenate(offs)).to(device),
                    torch.from_numpy(np.ascontiguousarray(flat, dtype)).to(device))

        t0 = env_tapes[0]
        volume = torch.from_numpy(np.stack([np.asarray(t.volume, np.int32) for t in env_tapes])).to(device)
        has_comp = getattr(t0, "comp_cents", None) is not None
        comp_off, comp = cat_csr("comp_off", "comp_cents", np.int32) if has_comp else (None, None)
        click_off, u_click = cat_csr("click_off", "u_click", np.float64)
        conv_off, u_conv = cat_csr("conv_off", "u_conv", np.float64)
        rev_off, rev = cat_csr("rev_off", "rev_cents", np.int32)
        impr = cost_off = cost = drift = None
        if getattr(t0, "impr", None) is not None:
            impr = torch.from_numpy(np.stack([np.asarray(t.impr, np.int32) for t in env_tapes])).to(device)
            cost_off, cost = cat_csr("cost_off", "cost", np.float64)
        if getattr(t0, "drift", None) is not None:
            drift = torch.from_numpy(np.stack([np.asarray(t.drift, np.float64) for t in env_tapes])).to(device)
        return DeviceTape(volume, comp_off, comp, click_off, u_click, conv_off, u_conv, rev_off, rev,
                          impr, cost_off, cost, drift)
