"""Per-call ("ephemeral") weights.

Three reference ops create fresh, unregistered parameters on the CPU inside every call and move
them to the input's device: cross_layer (DCN/dcn.py:37-45), residual_unit
(DeepCrossing/deepcrossing.py:37-39) and din_attention's att_net (DIN/din.py:61-67).  Parity
needs the same torch CPU-generator draws in the same order, so the draws stay in Python (in the
model files, with the reference's very constructor calls); this helper only ships the drawn
values to the GPU: one pinned staging tensor and one async copy into a device buffer whose
address stays fixed across steps (so a captured CUDA graph keeps reading the right place).
"""
from __future__ import annotations

import torch


class EphemeralBuffer:
    def __init__(self):
        self._dev = None

    def upload(self, tensors, device, fresh=False):
        """Copy CPU fp32 tensors to `device`; returns device views with the same shapes.
        `fresh=False`: into the buffer with the fixed address (what a captured CUDA graph reads; the
        previous draw is overwritten in place).  `fresh=True`: into a newly allocated device tensor,
        as the reference does with its per-call parameters — two forwards before a backward then
        keep two sets of weights, and autograd's saved-tensor version check guards them."""
        flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in tensors])
        if fresh:
            dev = torch.empty(flat.numel(), dtype=torch.float32, device=device)
        else:
            if self._dev is None or self._dev.numel() != flat.numel() or self._dev.device != device:
                self._dev = torch.empty(flat.numel(), dtype=torch.float32, device=device)
            dev = self._dev
        # a fresh pinned tensor per call: torch's host allocator keeps it alive until the copy ran
        dev.copy_(flat.pin_memory() if torch.cuda.is_available() else flat, non_blocking=True)
        return self._carve(dev, [t.shape for t in tensors])

    def views(self, shapes):
        return self._carve(self._dev, shapes)

    @staticmethod
    def _carve(dev, shapes):
        out, pos = [], 0
        for shp in shapes:
            n = 1
            for s in shp:
                n *= int(s)
            out.append(dev[pos:pos + n].view(*shp))
            pos += n
        return out

    @property
    def ready(self):
        return self._dev is not None
