"""Packed input staging: one pinned host buffer, one device buffer, ONE async copy per step.

The reference moves a batch with a `.to(device)` per tensor inside train() (26 copies per DIN
step: DIN/din.py:334-338; 7-9 for the other models, e.g. DCN/dcn.py:187-190).  `PackedBatch`
lays every tensor of a batch (nested dicts of int64 / float32 tensors, any shapes) out in one
byte buffer at 256-byte aligned offsets, once on the host (pinned) and once on the device; the
tensors the model sees are views into the device buffer, so

    packed = PackedBatch.like(example_batch, device)      # once
    packed.fill(batch)            # or let the loader collate straight into packed.host_views
    inputs = packed.to_device()   # one cudaMemcpyAsync on the current stream -> dict of device views

The device views keep their addresses from step to step, so they can be the static inputs of a
captured CUDA graph: a step is then one H2D copy + one graph launch.  Values are bit-identical to
the per-tensor path (bytes are copied, never converted; indices stay int64 as the reference's
torch.long).  SURVEY.md 8(f) "input staging".
"""
from __future__ import annotations

import torch

_ALIGN = 256


def _leaves(tree, prefix=()):
    if torch.is_tensor(tree):
        yield prefix, tree
    elif isinstance(tree, dict):
        for k, v in tree.items():
            yield from _leaves(v, prefix + (k,))
    else:
        raise TypeError(f"PackedBatch: unsupported leaf {type(tree).__name__} at {'/'.join(map(str, prefix))}")


def _build(paths_views):
    out = {}
    for path, view in paths_views:
        d = out
        for k in path[:-1]:
            d = d.setdefault(k, {})
        d[path[-1]] = view
    return out


class PackedBatch:
    def __init__(self, layout, nbytes, device, pin=True):
        """layout: list of (path, dtype, shape, byte offset)."""
        self.layout, self.nbytes = layout, nbytes
        self.device = torch.device(device)
        self.host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
        self.dev = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.host_views = _build(self._views(self.host))
        self.device_views = _build(self._views(self.dev))
        self._copied = None       # CUDA event recorded behind the last host -> device copy of `host`

    @classmethod
    def like(cls, batch, device, pin=True):
        layout, off = [], 0
        for path, t in _leaves(batch):
            if t.dtype not in (torch.int64, torch.float32, torch.int32, torch.bool, torch.uint8):
                raise TypeError(f"PackedBatch: dtype {t.dtype} of {'/'.join(map(str, path))} is not supported")
            layout.append((path, t.dtype, tuple(t.shape), off))
            off += (t.numel() * t.element_size() + _ALIGN - 1) // _ALIGN * _ALIGN
        return cls(layout, max(off, _ALIGN), device, pin)

    def _views(self, buf):
        for path, dtype, shape, off in self.layout:
            n = 1
            for s in shape:
                n *= s
            nb = n * torch.empty(0, dtype=dtype).element_size()
            yield path, buf[off:off + nb].view(dtype).view(shape)

    @property
    def payload_bytes(self):
        """Bytes of tensor data (without the alignment padding)."""
        total = 0
        for _, dtype, shape, _ in self.layout:
            n = 1
            for s in shape:
                n *= s
            total += n * torch.empty(0, dtype=dtype).element_size()
        return total

    def wait_host_free(self):
        """Block the host until the last asynchronous copy out of the pinned buffer has run: the copy
        of step k is queued behind step k-1's kernels, so a host that runs ahead of the GPU (any loop
        without a .item(), every CUDA-graph replay loop) would otherwise overwrite batch k with batch
        k+1 before the GPU has read it.  Call before writing `host` / `host_views` directly."""
        if self._copied is not None:
            self._copied.synchronize()
            self._copied = None

    def fill(self, batch):
        """Host-side collate into the pinned buffer (what a loader would do directly).  Waits for the
        previous to_device() copy of this buffer first (see wait_host_free)."""
        self.wait_host_free()
        leaves = dict(_leaves(batch))
        for path, dtype, shape, _ in self.layout:
            src = leaves[path]
            if src.dtype != dtype or tuple(src.shape) != shape:
                raise ValueError(f"PackedBatch.fill: {'/'.join(map(str, path))} is {src.dtype}{tuple(src.shape)}, "
                                 f"the layout holds {dtype}{shape}")
            dst = self.host_views
            for k in path:
                dst = dst[k]
            dst.copy_(src)
        return self

    def to_device(self, non_blocking=True):
        """One host-to-device copy of the whole batch on the current stream; returns the device views."""
        self.dev.copy_(self.host, non_blocking=non_blocking)
        if non_blocking and self.dev.is_cuda:
            self._copied = torch.cuda.Event()
            self._copied.record(torch.cuda.current_stream(self.dev.device))
        return self.device_views

    def load_from(self, other: "PackedBatch", non_blocking=True):
        """One copy of another packed batch (host or device side) into this one's device buffer."""
        if other.layout != self.layout:
            raise ValueError("PackedBatch.load_from: layouts differ")
        src = other.dev if other.dev.device == self.dev.device and other is not self else other.host
        self.dev.copy_(src, non_blocking=non_blocking)
        if src is other.host and non_blocking and self.dev.is_cuda:
            other._copied = torch.cuda.Event()
            other._copied.record(torch.cuda.current_stream(self.dev.device))
        return self.device_views


class Prefetcher:
    """Double-buffered input pipeline on top of PackedBatch: the host -> device copy of batch k+1 runs on a
    copy stream while step k computes; at the start of step k+1 one device -> device copy moves it into the
    step's static inputs (fixed addresses, so a captured CUDA graph keeps working).

        pf = Prefetcher(static)                   # static: the PackedBatch whose device views the model reads
        slot = pf.submit(host_batches[0])
        for k in range(steps):
            nxt = pf.submit(host_batches[k + 1])  # starts now, overlaps step k
            pf.consume(slot)                      # waits for ITS copy, then D2D into static.dev
            step()
            slot = nxt

    `after` (a CUDA event) delays a submitted copy until that point of another stream's timeline."""

    def __init__(self, static: PackedBatch):
        self.static = static
        dev = static.dev.device
        self.stage = [torch.empty_like(static.dev), torch.empty_like(static.dev)]
        self.stream = torch.cuda.Stream(device=dev)
        self.ready = [None, None]      # copy finished (recorded on the copy stream)
        self.freed = [None, None]      # stage buffer consumed (recorded on the consumer's stream)
        self.k = 0

    def submit(self, packed: PackedBatch, after=None) -> int:
        if packed.layout != self.static.layout:
            raise ValueError("Prefetcher.submit: layouts differ")
        slot = self.k % 2
        self.k += 1
        if after is not None:
            self.stream.wait_event(after)
        if self.freed[slot] is not None:
            self.stream.wait_event(self.freed[slot])
        with torch.cuda.stream(self.stream):
            self.stage[slot].copy_(packed.host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready[slot] = ev
        packed._copied = ev            # PackedBatch.fill waits for it before rewriting the pinned buffer
        return slot

    def wait(self, slot: int) -> None:
        """Make the current stream wait for the copy of `slot` (without consuming it)."""
        torch.cuda.current_stream(self.static.dev.device).wait_event(self.ready[slot])

    def consume(self, slot: int):
        main = torch.cuda.current_stream(self.static.dev.device)
        main.wait_event(self.ready[slot])
        self.static.dev.copy_(self.stage[slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        self.freed[slot] = ev
        return self.static.device_views
