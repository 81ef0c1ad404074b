"""Streamed inference over host batches: the caller-side loop of the hot path ("the callers either side").

The reference feeds HamGNN_pre / HamGNN_out from a DataLoader and copies every batch to the device inside the step
(`hamgnn/models/Model.py:128-179`, Lightning's `transfer_batch_to_device`); on a B200 the 1.2 GB of inputs and 1.1 GB of Hamiltonian
blocks of a tbg_m28 step take ~40 ms over PCIe when they run on the compute stream.  `streamed_forward` keeps the same per-batch
semantics (every batch is copied host -> device, every result device -> host) but puts the two copies on their own CUDA streams:
the inputs of batch i+1 and the result of batch i-1 travel while the kernels of batch i run.

    for h in streamed_forward(pre, out, host_batches):   # host_batches: pinned gd.Batch objects
        consume(h)                                       # h: pinned host tensor [N+E, nao^2], valid until two batches later
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional

import torch

from . import graph_data as gd
from . import lib as L


class _InputSlot:
    """One of the two device-side input buffer sets.  Batches of the same shapes are copied into the same tensors (no allocator
    traffic in steady state: a 1.2 GB cudaMalloc on the copy stream costs more than the copy it serves); `free` is the event
    after which the kernels of the slot's previous batch are done."""

    def __init__(self):
        self.tensors = {}
        self.free = None

    def fill(self, host: "gd.Batch", dev, s_in: "torch.cuda.Stream", compute: "torch.cuda.Stream") -> "gd.Batch":
        d = {}
        if self.free is not None:
            s_in.wait_event(self.free)
        for k, v in host.to_dict().items():
            if not torch.is_tensor(v):
                d[k] = v
                continue
            buf = self.tensors.get(k)
            if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                buf = torch.empty(v.shape, dtype=v.dtype, device=dev)   # allocated on the copy stream
                buf.record_stream(compute)                              # ... and read by the kernels on the compute stream
                self.tensors[k] = buf
            buf.copy_(v, non_blocking=True)
            d[k] = buf
        return gd.Batch(**d)


def streamed_forward(pre, out, host_batches: Iterable["gd.Batch"], device=None, key: str = "hamiltonian",
                     on_result: Optional[Callable[[int, torch.Tensor], None]] = None) -> Iterator[torch.Tensor]:
    """Yields the pinned host copy of `out(batch, pre(batch))[key]` for every batch of `host_batches`, in order.

    A yielded tensor is one of two pinned buffers per result shape and is overwritten two batches later; it is complete when it
    is yielded (the generator waits for that batch's device -> host copy, not for the device).  `on_result(i, host_tensor)` is
    called at the same point.  Inputs should be pinned (`Batch.pin_memory()`), otherwise the host -> device copies serialise."""
    dev = torch.device(device) if device is not None else next(pre.parameters()).device
    if dev.type != "cuda":
        raise L.HgbError("streamed_forward needs a CUDA device: the B200 kernels are the only implementation")
    compute = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    it = iter(host_batches)
    bufs = {}
    slots = [_InputSlot(), _InputSlot()]

    def stage(hb, idx):
        slot = slots[idx & 1]
        with torch.cuda.stream(s_in):
            b = slot.fill(hb, dev, s_in, compute)
            ev = torch.cuda.Event()
            ev.record(s_in)
        return b, ev, slot

    nxt = next(it, None)
    staged = stage(nxt, 0) if nxt is not None else None
    pending = None     # (index, host buffer, event) of the previous batch's result copy
    i = 0
    while staged is not None:
        b, ev_in, slot = staged
        nxt = next(it, None)
        staged = stage(nxt, i + 1) if nxt is not None else None      # inputs of batch i+1 travel while batch i computes
        compute.wait_event(ev_in)
        with torch.no_grad():
            res = out(b, pre(b))[key]
        ev_done = torch.cuda.Event()
        ev_done.record(compute)
        slot.free = ev_done                                           # the slot may be refilled (batch i+2) after these kernels
        shape = (tuple(res.shape), res.dtype)
        if shape not in bufs:
            bufs[shape] = [torch.empty(res.shape, dtype=res.dtype).pin_memory() for _ in range(2)]
        hbuf = bufs[shape][i & 1]
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done)
            hbuf.copy_(res, non_blocking=True)
            res.record_stream(s_out)
            ev_out = torch.cuda.Event()
            ev_out.record(s_out)
        if pending is not None:
            j, hprev, evp = pending
            evp.synchronize()
            if on_result is not None:
                on_result(j, hprev)
            yield hprev
        pending = (i, hbuf, ev_out)
        i += 1
    if pending is not None:
        j, hprev, evp = pending
        evp.synchronize()
        if on_result is not None:
            on_result(j, hprev)
        yield hprev
