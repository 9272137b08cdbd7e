"""Exchange steps of the sharded ClipLoss behind one small interface.

Three providers:
  * ``LocalComm``  world_size == 1, no exchange.
  * ``DistComm``   torch.distributed collectives (NCCL on GPUs, gloo in the CPU tests): all-gather,
                   all-reduce, reduce-scatter - the reference's own exchange steps (loss.py:32-38 and
                   the backward of torch.distributed.nn.all_gather).
  * ``NvlsComm``   single NVSwitch node: one symmetric-memory workspace
                   (torch.distributed._symmetric_memory) whose NVLink multicast alias is driven by the
                   multimem kernels of liboneprot_clip.so (oneprot_mc_*): the gather of the second
                   operand is ONE multimem.st pass (the switch replicates it), every reduction is a
                   multimem.ld_reduce pulled by its consumer (the switch adds), and the only
                   synchronisation is the 6-us symmetric-memory barrier.  ~4x less exchange time than
                   the NCCL calls at 8 GPUs for the 32768 x 1024 problem.

The interface speaks in the quantities of the algorithm, not in collectives:
  begin_forward(ops)            -> B_all, diag slot, sums views
  global_stats(stats_local)     -> stats identical on every rank (max)
  complete_sums()               -> [colsum | rowsum | diag] summed over ranks
  gather_grad_outputs(g32)      -> upstream gradients of all ranks
  db_buffer(N, d, dtype)        -> where the partial dB is written
  reduce_scatter_db(dbp)        -> this rank's rows of the summed dB
  sum_scalar(t)                 -> all-reduced scalar (d logit_scale)
"""
from __future__ import annotations

import os
import warnings
import weakref

import torch

try:
    import torch.distributed as dist
except ImportError:  # pragma: no cover
    dist = None


def _all_gather_rows(x, world_size, group=None):
    out = torch.empty((world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _reduce_scatter_rows(full, rank, world_size, group=None):
    n = full.shape[0] // world_size
    out = torch.empty((n,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM, group=group)
    else:  # gloo has no reduce-scatter
        tmp = full.contiguous().clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        out.copy_(tmp[rank * n:(rank + 1) * n])
    return out


class LocalComm:
    name = "local"

    def __init__(self, K):
        self.K = K

    def begin_forward(self, ops, rank, world):
        dev = ops.A.device
        N = ops.n
        sums = torch.zeros(3 * N, dtype=torch.float32, device=dev)
        stats = torch.zeros(4, dtype=torch.float32, device=dev)
        return dict(B_all=ops.B, sums=sums, stats=stats, stats_rows=ops.B, stats_off=0, token=None)

    def global_stats(self, st):
        return st["stats"]

    def complete_sums(self, st):
        return st["sums"]

    def gather_grad_outputs(self, g32, rank, world, token=None):
        return g32

    def b_all_for_backward(self, ops, B_all, token, rank, world, hold=None):
        return B_all

    def hold_for_backward(self, B_all):
        return None

    def db_buffer(self, N, d, dtype, dev):
        return torch.empty(N, d, dtype=dtype, device=dev)

    def reduce_scatter_db(self, dbp, rank, world, last_pass=True):
        return dbp

    def sum_scalar(self, t):
        return t

    def side_stream(self, dev):
        """Stream on which exchanges may overlap with compute (None: exchanges are synchronous)."""
        return None



class DistComm(LocalComm):
    """torch.distributed collectives; gathers first, so the row statistics see all of B at once."""
    name = "dist"

    def __init__(self, K, group):
        super().__init__(K)
        self.group = group

    def begin_forward(self, ops, rank, world):
        dev = ops.A.device
        N = world * ops.n
        B_all = _all_gather_rows(ops.B, world, self.group)
        sums = torch.zeros(3 * N, dtype=torch.float32, device=dev)
        stats = torch.zeros(4, dtype=torch.float32, device=dev)
        return dict(B_all=B_all, sums=sums, stats=stats, stats_rows=B_all, stats_off=rank * ops.n, token=None)

    def global_stats(self, st):
        dist.all_reduce(st["stats"], op=dist.ReduceOp.MAX, group=self.group)   # one reference G on every rank
        return st["stats"]

    def complete_sums(self, st):
        dist.all_reduce(st["sums"], op=dist.ReduceOp.SUM, group=self.group)
        return st["sums"]

    def gather_grad_outputs(self, g32, rank, world, token=None):
        return _all_gather_rows(g32, world, self.group)

    def reduce_scatter_db(self, dbp, rank, world, last_pass=True):
        return _reduce_scatter_rows(dbp, rank, world, self.group)

    def sum_scalar(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class NvlsComm(DistComm):
    """Symmetric-memory workspace + multimem kernels (one NVSwitch node).

    Workspace layout (bytes, one allocation, 256-byte aligned regions):
        Bg[2]     2 x cap_rows*W x dk bf16     gathered second operand, double buffered by call parity
        small[2]  2 x (3N + 4 + W4) fp32       [colsum | rowsum | diag], norm maxima, upstream grads
        dBp       N x d x 4                    partial dB (bf16 or fp32), reduced by its owner
    Double buffering + the barriers of the following call make reuse safe without extra
    synchronisation: a rank can only write buffer p of call t+2 after every rank entered call t+1,
    i.e. finished reading buffer p of call t (everything is stream-ordered).  That argument covers the
    BACKWARD of call t (it reads Bg[p] again: dL/dZ recompute and the dA GEMM) only when it is enqueued
    before forward t+1 - the reference's pattern (oneprot_module.py:100-105: one loss, its backward, the
    next modality).  For any other order (several forwards, then their backwards) begin_forward first
    snapshots the gathered operand of every forward that still waits for its backward
    (``hold_for_backward`` / ``_snapshot_pending``): at that point no peer can have overwritten it yet
    (it would have to be past forward t+1's barrier, which this rank has not entered).
    """
    name = "nvls"
    G_AT, STATS_AT, SUMS_AT = 0, 16, 32      # float offsets inside a small buffer: [g | maxima | sums]

    def __init__(self, K, group, world, rank, dev, static=False):
        super().__init__(K, group)
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.world, self.rank, self.dev = world, rank, dev
        # static = the CUDA-graph variant (graphed.py): every address, barrier channel and kernel argument of a step must
        # be the same at every replay, so there is ONE buffer set (no parity), no epoch-flagged fused gather, no side
        # stream, and a leading barrier per forward gives the write-after-read protection the double buffering gives
        # the eager provider (5 barriers per step instead of 3: meant for the launch-bound batch sizes)
        self.static = bool(static)
        self.cap = None          # (n, dk, d) capacity the workspace was sized for
        self.calls = 0
        self.gen = [0, 0]        # generation of the data held in Bg[p]
        self.chan = 0
        self._side = None
        self._pending = weakref.WeakSet()    # holders of forwards whose backward has not been enqueued yet

    # ---- workspace ----------------------------------------------------------------------
    @staticmethod
    def _al(x):
        return (x + 255) // 256 * 256

    def _ensure(self, n, dk, d):
        if self.cap is not None and n <= self.cap[0] and dk <= self.cap[1] and d <= self.cap[2]:
            return
        cap = (n, dk, d) if self.cap is None else (max(n, self.cap[0]), max(dk, self.cap[1]), max(d, self.cap[2]))
        W = self.world
        N = W * cap[0]
        self.bg_bytes = self._al(N * cap[1] * 2)
        self.small_floats = (self.SUMS_AT + 3 * N + 63) // 64 * 64
        self.small_bytes = self.small_floats * 4
        self.db_bytes = self._al(N * cap[2] * 4)
        self.ctl_bytes = 4096           # flags u32[W][9] @0, stats_all f32[W][4] @1024, counters u32[8] @2048
        total = 2 * self.bg_bytes + 2 * self.small_bytes + self.db_bytes + self.ctl_bytes
        buf = self.symm.empty(total, dtype=torch.uint8, device=self.dev)
        self.hdl = self.symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        if not self.hdl.multicast_ptr:
            raise RuntimeError("symmetric memory has no multicast (NVLS) mapping on this system")
        self.buf = buf
        self.mc = int(self.hdl.multicast_ptr)
        self.cap = cap
        self.gen = [0, 0]            # tokens of earlier forwards no longer match: they re-gather
        self.buf.zero_()
        self._barrier()

    def _barrier(self):
        self.hdl.barrier(self.chan)
        self.chan = (self.chan + 1) % 8

    def side_stream(self, dev):
        if self.static:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    def _view(self, off, shape, dtype):
        nbytes = dtype.itemsize
        numel = 1
        for s in shape:
            numel *= s
        return self.buf[off:off + numel * nbytes].view(dtype).view(*shape)

    def _off_bg(self, p):
        return p * self.bg_bytes

    def _off_small(self, p):
        return 2 * self.bg_bytes + p * self.small_bytes

    def _off_db(self):
        return 2 * self.bg_bytes + 2 * self.small_bytes

    def _off_ctl(self):
        return 2 * self.bg_bytes + 2 * self.small_bytes + self.db_bytes

    # ---- gathered operand of forwards that still wait for their backward -----------------
    class _Hold:
        """Owned by the autograd ctx of one forward (dies with it); view = the rows of Bg[p] it used."""
        __slots__ = ("view", "snap", "__weakref__")

    def hold_for_backward(self, B_all):
        if self.static:
            return None               # a graphed step refuses a backward that is not the latest forward's (graphed.py)
        h = NvlsComm._Hold()
        h.view, h.snap = B_all, None
        self._pending.add(h)
        return h

    def _snapshot_pending(self):
        for h in list(self._pending):
            if h.snap is None:
                h.snap = h.view.clone()       # stream-ordered before this call's exchanges

    # ---- forward -------------------------------------------------------------------------
    def begin_forward(self, ops, rank, world, fused=True):
        """fused=False: always gather with the separate multicast kernel (callers whose forward kernel does not
        carry the all-gather: SigLipLoss)."""
        K = self.K
        n, dk = ops.n, ops.dk
        N = world * n
        self._ensure(n, dk, ops.d)
        if self.static:
            fused = False
            self._barrier()           # every rank has finished the previous step's reads of the single buffer set
            p, self.gen[0] = 0, 1
            self.calls = 1
        else:
            self._snapshot_pending()
            p = self.calls & 1
            self.calls += 1
            self.gen[p] = self.calls
        Bg = self._view(self._off_bg(p), (N, dk), torch.bfloat16)
        small = self._view(self._off_small(p), (self.SUMS_AT + 3 * N,), torch.float32)
        small.zero_()
        st = dict(B_all=Bg, sums=small[self.SUMS_AT:self.SUMS_AT + 3 * N], stats=small[self.STATS_AT:self.STATS_AT + 4],
                  stats_rows=ops.B, stats_off=0, token=(p, self.calls), p=p, N=N, ag=None)
        chunks = next((c for c in (8, 4, 2, 1) if n % (c * 256) == 0), 0) if fused else 0
        if chunks:
            # the gather is fused into the forward kernel: nothing to launch here
            base = int(self.hdl.buffer_ptrs[rank]) + self._off_ctl()
            mcb = self.mc + self._off_ctl()
            st["stats_glob"] = torch.zeros(4, dtype=torch.float32, device=self.dev)
            st["ag"] = dict(src=ops.B.data_ptr(), dst_mc=self.mc + self._off_bg(p) + rank * n * dk * 2,
                            counters=base + 2048, flags_mc=mcb, flags=base, stats_mc=mcb + 1024, stats_all=base + 1024,
                            stats_out=st["stats_glob"].data_ptr(), epoch=self.calls & 0x7fffffff, rank=rank, world=world,
                            chunks=chunks, rows_per_rank=n)
        else:
            # second operand: one multimem.st pass puts this rank's rows into every GPU's Bg[p]
            K.mc_store(ops.B, self.mc + self._off_bg(p) + rank * n * dk * 2, n * dk * 2)
        return st

    def gather_rows(self, ops, rank, world):
        """All-gather of the second operand alone (one multimem.st pass + barrier) -> (B_all view, token)."""
        st = self.begin_forward(ops, rank, world, fused=False)
        self._barrier()                           # every rank's rows are in place
        return st["B_all"], st["token"]

    def global_stats(self, st):
        # (the caller ran rowstats on the LOCAL rows: stats holds this rank's maxima)
        if st.get("ag") is not None:
            return st["stats"]                    # published and maximised inside the forward kernel
        self._barrier()                           # B rows + local maxima of every rank are in place
        out = torch.empty(4, dtype=torch.float32, device=self.dev)
        self.K.mc_allreduce_f32(self.mc + self._off_small(st["p"]) + self.STATS_AT * 4, out, 4, 1)
        return out

    def complete_sums(self, st):
        self._barrier()
        N = st["N"]
        cnt = (3 * N + 3) // 4 * 4
        out = torch.empty(cnt, dtype=torch.float32, device=self.dev)
        self.K.mc_allreduce_f32(self.mc + self._off_small(st["p"]) + self.SUMS_AT * 4, out, cnt, 0)
        return out[0:3 * N]

    # ---- raw addresses for the host-side step sequencer (sequencer.py) -----------------------
    def seq_ready(self, n, dev):
        """The sequencer covers the fused-gather forward and the side-stream backward only."""
        chunks = next((c for c in (8, 4, 2, 1) if n % (c * 256) == 0), 0)
        return bool(chunks) and self.side_stream(dev) is not None

    def _base(self):
        return int(self.hdl.buffer_ptrs[self.rank])

    def seq_forward_desc(self, ops, rank, world, stats_out_addr):
        """begin_forward without its memsets and views: addresses only (the C side clears the
        [g | maxima | sums] buffer itself)."""
        n, dk = ops.n, ops.dk
        N = world * n
        self._ensure(n, dk, ops.d)
        self._snapshot_pending()
        p = self.calls & 1
        self.calls += 1
        self.gen[p] = self.calls
        base, small = self._base(), self._off_small(p)
        ctl, mcb = base + self._off_ctl(), self.mc + self._off_ctl()
        chunks = next(c for c in (8, 4, 2, 1) if n % (c * 256) == 0)
        # field order of oneprot_ag_t
        ag = (ops.B.data_ptr(), self.mc + self._off_bg(p) + rank * n * dk * 2, ctl + 2048, mcb, ctl, mcb + 1024, ctl + 1024,
              stats_out_addr, self.calls & 0x7fffffff, rank, world, chunks, n)
        return dict(B_all=base + self._off_bg(p), B_view=self._view(self._off_bg(p), (N, dk), torch.bfloat16),
                    stats=base + small + self.STATS_AT * 4, zero_ptr=base + small,
                    zero_bytes=(self.SUMS_AT + 3 * N) * 4, sums=base + small + self.SUMS_AT * 4,
                    sums_mc=self.mc + small + self.SUMS_AT * 4, ag=ag, token=(p, self.calls))

    def seq_backward_desc(self, n, d, rank, token):
        base, g_off = self._base(), self._off_small(token[0]) + self.G_AT * 4
        return dict(dB=base + self._off_db(), dB_mc_mine=self.mc + self._off_db() + rank * n * d * 2,
                    g_slot=base + g_off, g_slot_mc=self.mc + g_off)

    def seq_handle(self):
        """Event holder of the C sequencer (one per provider)."""
        if getattr(self, "_seq", None) is None:
            import ctypes
            from . import _lib
            h = ctypes.c_void_p()
            _lib.check(_lib.load().oneprot_seq_create(ctypes.byref(h)), "oneprot_seq_create")
            self._seq = h
        return self._seq

    # ---- backward ------------------------------------------------------------------------
    def gather_grad_outputs(self, g32, rank, world, token=None):
        p = token[0] if token is not None else ((self.calls - 1) & 1)
        W4 = (world + 3) // 4 * 4
        off = self._off_small(p) + self.G_AT * 4
        slot = self._view(off, (W4,), torch.float32)
        slot.zero_()
        slot[rank:rank + 1].copy_(g32)
        self._barrier()
        out = torch.empty(W4, dtype=torch.float32, device=self.dev)
        self.K.mc_allreduce_f32(self.mc + off, out, W4, 0)   # one-hot contributions: the sum is the gather
        return out[0:world]

    def b_all_for_backward(self, ops, B_all, token, rank, world, hold=None):
        """The gathered operand for the backward of the forward `token`: the symmetric buffer when that backward
        directly follows its forward, else the snapshot begin_forward took (see the class docstring)."""
        if self.static:
            return B_all
        if hold is not None:
            self._pending.discard(hold)
            if hold.snap is not None:
                return hold.snap
        p, gen = token
        if self.gen[p] != gen:       # no holder (workspace regrown, or a caller without one): gather again
            return _all_gather_rows(ops.B, world, self.group)
        return B_all

    def db_buffer(self, N, d, dtype, dev):
        return self._view(self._off_db(), (N, d), dtype)

    def reduce_scatter_db(self, dbp, rank, world, last_pass=True):
        K = self.K
        N, d = dbp.shape
        n = N // world
        self._barrier()                           # every rank's partial dB is written
        out = torch.empty(n, d, dtype=dbp.dtype, device=dbp.device)
        esz = dbp.element_size()
        src = self.mc + self._off_db() + rank * n * d * esz
        if dbp.dtype == torch.bfloat16:
            K.mc_reduce_bf16(src, out, n * d * esz)
        else:
            K.mc_allreduce_f32(src, out, n * d, 0)
        if not last_pass:
            self._barrier()                       # the next pass must not overwrite dBp before all owners pulled
        return out


def make_comm(K, world, rank, group, device, prefer=None):
    """Pick the exchange provider.  ONEPROT_COMM = auto | dist | nvls overrides."""
    if world == 1:
        return LocalComm(K)
    mode = (prefer or os.environ.get("ONEPROT_COMM", "auto")).lower()
    if mode.startswith("nvls-static"):   # "nvls-static/<shape tag>": one workspace per graphed step shape; CUDA-graph replay (graphed.py): no fallback, the torch.distributed provider is not captured
        return NvlsComm(K, group, world, rank, device, static=True)
    want_nvls = mode == "nvls" or (mode == "auto" and device.type == "cuda" and dist.get_backend(group) == "nccl"
                                   and hasattr(K, "mc_store") and world <= torch.cuda.device_count())
    if want_nvls:
        try:
            c = NvlsComm(K, group, world, rank, device)
            return c
        except Exception as e:  # pragma: no cover
            if mode == "nvls":
                raise
            warnings.warn(f"oneprot_b200: NVLS exchange unavailable ({e!r}); using torch.distributed collectives")
    return DistComm(K, group)
