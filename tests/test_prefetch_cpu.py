"""CPU: stream-ordering logic of PinnedPairPrefetcher (oneprot_b200/prefetch.py) under a vector-clock
model of CUDA streams and events.  Streams, events and the copies are stand-ins; what is checked is
the happens-before relation the real streams would enforce:
  (1) a consumer read of a slot happens after the copy that filled it, and
  (2) the copy that overwrites a slot happens after everything the consumer queued while it owned
      the slot (i.e. up to the following next()),
for the next -> submit -> compute pattern of bench.py and for bursts of submits, with 2 and 3 slots."""
import contextlib

import pytest
import torch

from oneprot_b200 import prefetch as pf_mod


class VStream:
    """A stream = a sequence of operations; its clock is a dict stream-id -> ops known to have completed."""
    n = 0

    def __init__(self, device=None):
        VStream.n += 1
        self.id = VStream.n
        self.clock = {self.id: 0}
        self.log = []

    def tick(self, what):
        self.clock[self.id] += 1
        self.log.append((what, dict(self.clock)))
        return dict(self.clock)

    def record_event(self):
        return VEvent(self.tick("record"))

    def wait_event(self, ev):
        for k, v in ev.clock.items():
            self.clock[k] = max(self.clock.get(k, 0), v)
        self.tick("wait")

    def wait_stream(self, other):
        self.wait_event(VEvent(dict(other.clock)))


class VEvent:
    def __init__(self, clock):
        self.clock = clock


def happens_before(a, b):
    """clock a (taken when an op was enqueued on its stream) is covered by clock b"""
    return all(b.get(k, 0) >= v for k, v in a.items())


class Model:
    def __init__(self, monkeypatch):
        self.consumer = VStream()
        self.stack = [self.consumer]
        self.copies = []     # (slot buffer ptr, clock of the copy op)
        monkeypatch.setattr(torch.cuda, "Stream", VStream)
        monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: self.stack[-1])
        monkeypatch.setattr(torch.cuda, "stream", self.use)
        monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
        monkeypatch.setattr(torch.Tensor, "is_pinned", lambda t: True)
        model = self
        orig_copy = torch.Tensor.copy_

        def copy_(dst, src, non_blocking=False):
            model.copies.append((dst.data_ptr(), model.stack[-1].tick("copy")))
            return orig_copy(dst, src)
        monkeypatch.setattr(torch.Tensor, "copy_", copy_)
        orig_empty = torch.empty
        monkeypatch.setattr(torch, "empty", lambda *a, device=None, **k: orig_empty(*a, **k))
        monkeypatch.setattr(pf_mod.torch, "device", lambda d: type("D", (), {"type": "cuda"})())

    @contextlib.contextmanager
    def use(self, s):
        self.stack.append(s)
        try:
            yield
        finally:
            self.stack.pop()


def _pairs(k, n=4, d=8):
    return [(torch.full((n, d), float(i)), torch.full((n, d), float(-i))) for i in range(k)]


@pytest.mark.parametrize("slots", [2, 3])
def test_next_submit_compute_pattern_is_race_free(monkeypatch, slots):
    m = Model(monkeypatch)
    pf = pf_mod.PinnedPairPrefetcher("cuda", slots=slots)
    pairs = _pairs(9)
    pf.submit(*pairs[0])
    reads = []       # (slot ptr, clock at next(), clock after the compute that used the pair, expected value)
    for k in range(len(pairs)):
        A, B = pf.next()
        start = dict(m.consumer.clock)
        assert float(A[0, 0]) == float(k) and float(B[0, 0]) == float(-k)     # the right pair, in order
        if k + 1 < len(pairs):
            pf.submit(*pairs[k + 1])
        end = m.consumer.tick("compute")                                     # kernels reading A, B
        reads.append((A.data_ptr(), start, end))
    a_copies = [(p, c) for p, c in m.copies if any(p == r[0] for r in reads)]
    assert len(a_copies) == len(pairs)
    for k, (ptr, start, end) in enumerate(reads):
        fill = a_copies[k]
        assert fill[0] == ptr
        assert happens_before(fill[1], start), f"pair {k} read before its copy landed"
        for j in range(k + 1, len(pairs)):                                   # later copies into the same slot
            if a_copies[j][0] == ptr:
                assert happens_before(end, a_copies[j][1]), f"copy {j} may overwrite pair {k} while it is in use"
                break
    # and the copy of pair k+1 does NOT wait for the compute of pair k (that is the overlap)
    assert not happens_before(reads[3][2], a_copies[4][1])


def test_burst_of_submits_and_ring_errors(monkeypatch):
    m = Model(monkeypatch)
    pf = pf_mod.PinnedPairPrefetcher("cuda", slots=3)
    pairs = _pairs(6)
    for p in pairs[:3]:
        pf.submit(*p)
    with pytest.raises(RuntimeError):
        pf.submit(*pairs[3])                  # ring full
    A0, _ = pf.next()
    with pytest.raises(RuntimeError):
        pf.submit(*pairs[3])                  # slot 0 is owned by the consumer until the following next()
    m.consumer.tick("compute")
    A1, _ = pf.next()
    pf.submit(*pairs[3])                      # overwrites slot 0 after the release recorded by this next()
    assert float(A1[0, 0]) == 1.0
    A2, _ = pf.next()
    A3, _ = pf.next()
    assert float(A2[0, 0]) == 2.0 and float(A3[0, 0]) == 3.0 and A3.data_ptr() == A0.data_ptr()
    with pytest.raises(RuntimeError):
        pf.next()                             # nothing left
