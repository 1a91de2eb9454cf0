"""Replay tapes on the device (the `adc_tape` of include/adcraft_b200.h).

A tape holds, for every (env, keyword) unit of ONE env step, the pre-drawn values in the order
the reference consumes them: the volume, one competitor bid per auction, one
uniform per click slot, one per accepted click, one revenue per conversion; for explicit
keywords the per-sub-step impression counts and one cost per impression; optional drift
coefficients.  Streams are CSR over the E*K units.
"""
from __future__ import annotations

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
    packed: Optional[torch.Tensor] = None       # uint8, one 16-byte aligned record per unit
    packed_off: Optional[torch.Tensor] = None   # int64 [E*K+1] byte offsets
    comp_f64: Optional[torch.Tensor] = None     # multi-bidder keywords: clearing price per auction

    def c_struct(self) -> _capi.Tape:
        t = _capi.Tape()
        for name, _ in _capi.Tape._fields_:
            v = getattr(self, name)
            setattr(t, name, None if v is None else v.data_ptr())
        return t

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.__dict__.values() if v is not None)

    def pack(self) -> "DeviceTape":
        """Build the packed copy of an implicit-keyword tape (include/adcraft_b200.h, `adc_tape`):
        per unit one record  hdr[8] (V, counts, flags) | comp (padded to 4 entries with INT32_MAX) | click | conv | rev
        at a 16-byte aligned offset, so the replay kernel fetches a unit's day with one bulk copy.
        Pure tensor ops on whatever device the tape lives on; returns self."""
        if self.comp_cents is None:
            raise ValueError("pack(): only implicit-keyword tapes have a packed form")
        dev = self.volume.device
        i64 = torch.int64
        V = self.volume.reshape(-1).to(i64)
        U = V.numel()
        live = V > 0

        def lens(off):
            return torch.where(live, off[1:] - off[:-1], torch.zeros_like(V))

        n_comp = torch.minimum(lens(self.comp_off), V)
        n_click, n_conv, n_rev = lens(self.click_off), lens(self.conv_off), lens(self.rev_off)
        comp_pad = (n_comp + 3) & ~3
        body = 32 + 4 * comp_pad + 8 * n_click + 8 * n_conv + 4 * n_rev
        size = torch.where(live, (body + 15) & ~15, torch.zeros_like(V))
        if int(size.max()) > 0x7FFFFFF0:
            raise ValueError("pack(): a unit's record exceeds 2 GiB")
        off = torch.zeros(U + 1, dtype=i64, device=dev)
        off[1:] = torch.cumsum(size, 0)
        buf = torch.zeros(max(int(off[-1]), 16), dtype=torch.uint8, device=dev)
        b32, b64 = buf.view(torch.int32), buf.view(torch.float64)
        rec = off[:-1]
        h = (rec[live] // 4)
        for j, col in enumerate((V, n_comp, n_click, n_conv, n_rev)):
            b32[h + j] = col[live].to(torch.int32)

        def scatter(dst, dst_base, counts, src=None, src_base=None, fill=None):
            total = int(counts.sum())
            if total == 0:
                return
            unit = torch.repeat_interleave(torch.arange(U, device=dev), counts)
            k = torch.arange(total, device=dev) - (torch.cumsum(counts, 0) - counts)[unit]
            if src is None:
                dst[dst_base[unit] + k] = fill
            else:
                dst[dst_base[unit] + k] = src[src_base[unit] + k]

        comp_base = (rec + 32) // 4
        scatter(b32, comp_base + n_comp, comp_pad - n_comp, fill=0x7FFFFFFF)
        scatter(b32, comp_base, n_comp, self.comp_cents, self.comp_off[:-1])
        click_base = (rec + 32 + 4 * comp_pad) // 8
        scatter(b64, click_base, n_click, self.u_click, self.click_off[:-1])
        scatter(b64, click_base + n_click, n_conv, self.u_conv, self.conv_off[:-1])
        scatter(b32, (click_base + n_click + n_conv) * 2, n_rev, self.rev_cents, self.rev_off[:-1])
        # flags (hdr[5]) bit 0, ADC_PACKED_NARROW: no negative competitor bid, revenues in [0, 65535]
        def seg_extreme(vals, base, counts, reduce, init):
            out = torch.full((U,), init, dtype=i64, device=dev)
            total = int(counts.sum())
            if total:
                unit = torch.repeat_interleave(torch.arange(U, device=dev), counts)
                k = torch.arange(total, device=dev) - (torch.cumsum(counts, 0) - counts)[unit]
                out.scatter_reduce_(0, unit, vals[base[unit] + k].to(i64), reduce, include_self=True)
            return out

        comp_min = seg_extreme(self.comp_cents, self.comp_off[:-1], n_comp, "amin", 0)
        rev_min = seg_extreme(self.rev_cents, self.rev_off[:-1], n_rev, "amin", 0)
        rev_max = seg_extreme(self.rev_cents, self.rev_off[:-1], n_rev, "amax", 0)
        narrow = (comp_min >= 0) & (rev_min >= 0) & (rev_max <= 65535)
        b32[h + 5] = narrow[live].to(torch.int32)
        self.packed, self.packed_off = buf, off
        return self

    def trimmed(self, impressions, clicks, conversions) -> "DeviceTape":
        """The tape a recording of this very step would hold: every stream cut to what the step
        consumed (comp: the volume, click: impressions, conv: clicks, rev: conversions).  Streams
        of a loose synthetic tape are longer; replaying either gives identical outcomes."""
        dev = self.volume.device
        V = self.volume.reshape(-1).to(torch.int64)
        U = V.numel()

        def cut(off, vals, counts):
            counts = torch.minimum(counts.reshape(-1).to(torch.int64), off[1:] - off[:-1])
            new_off = torch.zeros(U + 1, dtype=torch.int64, device=dev)
            new_off[1:] = torch.cumsum(counts, 0)
            total = int(new_off[-1])
            unit = torch.repeat_interleave(torch.arange(U, device=dev), counts)
            k = torch.arange(total, device=dev) - new_off[:-1][unit]
            out = vals[off[:-1][unit] + k] if total else vals[:1].clone()
            return new_off, out

        comp_off, comp = cut(self.comp_off, self.comp_cents, V)
        click_off, click = cut(self.click_off, self.u_click, impressions)
        conv_off, conv = cut(self.conv_off, self.u_conv, clicks)
        rev_off, rev = cut(self.rev_off, self.rev_cents, conversions)
        return DeviceTape(self.volume, comp_off, comp, click_off, click, conv_off, conv, rev_off, rev,
                          drift=self.drift)

    @staticmethod
    def from_host(env_tapes: Sequence[object], device, pack: bool = False) -> "DeviceTape":
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
        impr = cost_off = cost = drift = comp_f64 = None
        if getattr(t0, "impr", None) is not None:
            impr = torch.from_numpy(np.stack([np.asarray(t.impr, np.int32) for t in env_tapes])).to(device)
            if getattr(t0, "cost", None) is not None:
                cost_off, cost = cat_csr("cost_off", "cost", np.float64)
        if getattr(t0, "comp_f64", None) is not None:  # multi-bidder keywords: highest bid per auction
            _, comp_f64 = cat_csr("comp_off", "comp_f64", np.float64)
        if getattr(t0, "drift", None) is not None:
            drift = torch.from_numpy(np.stack([np.asarray(t.drift, np.float64) for t in env_tapes])).to(device)
        tape = DeviceTape(volume, comp_off, comp, click_off, u_click, conv_off, u_conv, rev_off, rev,
                          impr, cost_off, cost, drift, comp_f64=comp_f64)
        return tape.pack() if pack and has_comp and comp_f64 is None else tape
