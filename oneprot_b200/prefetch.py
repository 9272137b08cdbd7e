"""Host -> device staging of the (anchor, modality) embedding pair in front of ``ClipLoss``.

The reference's loss receives device tensors from the encoders (``oneprot_module.py:93-100``); when
the embeddings come from HOST memory instead (pre-computed embedding shards, ``bench.py``'s
end-to-end leg) the copy of 2 n d elements is as long as a third of the fused fwd+bwd at
n = 32768 (128 MiB over PCIe gen5 = 2.4 ms against 6.4 ms).  ``PinnedPairPrefetcher`` hides it:
the pair of step k+1 is copied on its own CUDA stream into the second of two device slots while
step k computes.  Plumbing only - torch streams, events and ``copy_``; no arithmetic.

    pf = PinnedPairPrefetcher(device)
    pf.submit(a_host, b_host)                 # pinned CPU tensors
    for ...:
        A, B = pf.next()                      # device tensors, valid on the current stream
        pf.submit(next_a_host, next_b_host)   # flies under the compute below
        loss = clip(A.requires_grad_(), B.requires_grad_()); loss.backward()
"""
from __future__ import annotations

import torch


class PinnedPairPrefetcher:
    """Ring of ``slots`` device buffers filled by H2D copies on a dedicated copy stream.

    ``submit`` enqueues the copy of one pair (pinned host tensors) into the next slot of the ring
    and never blocks the host.  ``next`` makes the current stream wait for the oldest submitted
    pair and returns detached views of its slot; the pair stays valid until the FOLLOWING call of
    ``next`` (with the default two slots), at which point an event recorded on the consumer's stream
    releases the slot: the copy that overwrites it waits for that event on the copy stream.  Hence
    the calling pattern  next -> submit -> compute  keeps one copy in flight under every compute."""

    def __init__(self, device, slots: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("PinnedPairPrefetcher stages into CUDA memory; got device " + str(device))
        if slots < 2:
            raise ValueError("need at least two slots (one consumed, one in flight)")
        self.slots = slots
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._bufs = [None] * slots       # (A_dev, B_dev) per slot, allocated on first use
        self._done = [None] * slots       # copy-stream event: the slot's pair has landed
        self._release = [None] * slots    # consumer-stream event: the slot may be overwritten after it
        self._submitted = 0
        self._consumed = 0

    def _slot_buffers(self, slot, a_host, b_host):
        """Device buffers of a slot; (re)allocated when the pair's shape or dtype changes."""
        cur = self._bufs[slot]
        if (cur is not None and cur[0].shape == a_host.shape and cur[0].dtype == a_host.dtype
                and cur[1].shape == b_host.shape and cur[1].dtype == b_host.dtype):
            return cur, False
        cur = (torch.empty(a_host.shape, dtype=a_host.dtype, device=self.device),
               torch.empty(b_host.shape, dtype=b_host.dtype, device=self.device))
        self._bufs[slot] = cur
        return cur, True

    def submit(self, a_host: torch.Tensor, b_host: torch.Tensor) -> None:
        if a_host.is_cuda or b_host.is_cuda:
            raise ValueError("submit() takes host tensors")
        if not (a_host.is_pinned() and b_host.is_pinned()):
            raise ValueError("submit() needs pinned host tensors (pageable memory makes the copy synchronous)")
        w = self._submitted
        slot = w % self.slots
        with torch.cuda.device(self.device):
            if w >= self.slots:
                # the slot held pair w - slots, released by the next() call after the one that returned it
                if self._consumed < w - self.slots + 2:
                    raise RuntimeError("prefetcher is full: call next() before submitting another pair")
                self.copy_stream.wait_event(self._release[slot])
            (A_dev, B_dev), fresh = self._slot_buffers(slot, a_host, b_host)
            if fresh:
                # the caching allocator handed the block out in the consumer stream's order
                self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.copy_stream):
                A_dev.copy_(a_host, non_blocking=True)
                B_dev.copy_(b_host, non_blocking=True)
                self._done[slot] = self.copy_stream.record_event()
        self._submitted = w + 1

    def next(self):
        """(A, B) of the oldest submitted pair as detached device tensors; the current stream waits
        for their copy.  The pair returned by the previous call is released at this point of the
        current stream."""
        r = self._consumed
        if r >= self._submitted:
            raise RuntimeError("next() without a submitted pair")
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            if r >= 1:
                self._release[(r - 1) % self.slots] = cur.record_event()
            slot = r % self.slots
            cur.wait_event(self._done[slot])
        self._consumed = r + 1
        A_dev, B_dev = self._bufs[slot]
        return A_dev.detach(), B_dev.detach()
