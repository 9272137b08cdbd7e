"""CPU: the host-side step sequencer (oneprot_b200/csrc/clip_sequence.cu + sequencer.py) enqueues
exactly what the Python host enqueues.

Both paths run in the library's DRY trace mode (oneprot_trace_begin(1): every entry point records
its arguments and returns before any CUDA call), on CPU tensors whose addresses only label the
trace.  Streams, events and the NVLS provider's symmetric memory are stand-ins that write their own
trace lines.  The two traces must agree line by line after canonicalisation:
  * addresses inside buffers both paths share (A, B, logit_scale, the symmetric workspace and its
    multicast alias) are compared as region + exact byte offset,
  * every other address (temporaries: separate torch tensors on the Python path, slices of one
    workspace on the C path) is renamed by order of first appearance, which checks that the same
    buffer flows between the same producers and consumers,
  * memsets / copies the Python path performs with torch ops are checked against the C trace
    explicitly."""
import contextlib
import re

import pytest
import torch

from oneprot_b200 import clip_loss as cl
from oneprot_b200 import comm as comm_mod
from oneprot_b200 import kernels as K

MAIN, SIDE = 0x1000, 0x2000
MC_BASE = 0x7F0000000000


# ---------------------------------------------------------------------------------------------
# stand-ins for CUDA streams / events / symmetric memory
# ---------------------------------------------------------------------------------------------
class FakeEvent:
    count = 0

    def __init__(self):
        FakeEvent.count += 1
        self.id = 100 + FakeEvent.count


class FakeStream:
    def __init__(self, handle):
        self.cuda_stream = handle

    def record_event(self):
        ev = FakeEvent()
        K.trace_note(f"record ev={ev.id} st={self.cuda_stream:#x}")
        return ev

    def wait_event(self, ev):
        K.trace_note(f"wait ev={ev.id} st={self.cuda_stream:#x}")


class Streams:
    def __init__(self):
        self.main, self.side = FakeStream(MAIN), FakeStream(SIDE)
        self.stack = [self.main]

    def current(self, *a, **k):
        return self.stack[-1]

    @contextlib.contextmanager
    def use(self, s):
        self.stack.append(s)
        try:
            yield
        finally:
            self.stack.pop()


class FakeHandle:
    def __init__(self, buf, world, streams):
        self.buffer_ptrs = [buf.data_ptr()] * world
        self.multicast_ptr = MC_BASE
        self.streams = streams

    def barrier(self, channel):
        K.trace_note(f"barrier st={self.streams.current().cuda_stream:#x}")


class FakeSymm:
    def __init__(self, world, streams):
        self.world, self.streams = world, streams

    def empty(self, total, dtype, device):
        return torch.zeros(total, dtype=dtype)

    def rendezvous(self, buf, group):
        return FakeHandle(buf, self.world, self.streams)


class FakeNvlsComm(comm_mod.NvlsComm):
    """The real provider logic over CPU memory: only the symmetric allocation, the barrier and the
    side stream are stand-ins."""

    def __init__(self, world, rank, streams):
        super().__init__(K, object(), world, rank, torch.device("cpu"))
        self.symm = FakeSymm(world, streams)
        self.streams = streams

    def side_stream(self, dev):
        return self.streams.side


@pytest.fixture
def streams(monkeypatch):
    s = Streams()
    # keep every temporary of a traced step alive: a freed block that the allocator hands out again
    # would give two different buffers the same name in the canonical trace
    keep = []
    for fn_name in ("empty", "zeros", "full"):
        orig = getattr(torch, fn_name)

        def keeping(*a, _orig=orig, **k):
            t = _orig(*a, **k)
            keep.append(t)
            return t
        monkeypatch.setattr(torch, fn_name, keeping)
    s.keep = keep
    monkeypatch.setattr(torch.cuda, "current_stream", s.current)
    monkeypatch.setattr(torch.cuda, "stream", s.use)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, st: None)
    return s


# ---------------------------------------------------------------------------------------------
# running one fwd+bwd under the dry trace
# ---------------------------------------------------------------------------------------------
def _cfg(world, rank, local_loss, gwg, panel_bytes, seq, keep_exp=False):
    return dict(world_size=world, rank=rank, group=None, local_loss=local_loss, gather_with_grad=gwg, loss_dtype=torch.float32,
                panel_bytes=panel_bytes, host_sequencer=seq, keep_exp=keep_exp)


def _run(streams, A, B, scale, cfg, comm, need=(True, True)):
    """-> trace lines of one forward + backward through _ClipLossFunction."""
    get_comm = cl._get_comm
    cl._get_comm = lambda *a, **k: comm          # the exchange provider under test
    try:
        a = A.clone().requires_grad_(need[0])
        b = B.clone().requires_grad_(need[1])
        with K.launch_trace(dry_stream=lambda: streams.current().cuda_stream) as tr:
            loss, _, _ = cl._ClipLossFunction.apply(a, b, scale, cfg)
            K.trace_note("---- backward")
            loss.backward()
        regions = [("A", a), ("B", b), ("scale", scale)]
        return tr.lines, [(nm, t.data_ptr(), t.numel() * t.element_size()) for nm, t in regions]
    finally:
        cl._get_comm = get_comm


PTR = re.compile(r"=(0x[0-9a-f]+|\(nil\))")


def _canon(lines, regions, drop=("memset", "copy")):
    names, ev_names, out = {}, {}, []
    for ln in lines:
        if ln.split()[0] in drop:
            continue

        def sub(m):
            tok = m.group(1)
            if tok == "(nil)":
                return "=null"
            v = int(tok, 16)
            if v in (MAIN, SIDE):
                return "=main" if v == MAIN else "=side"
            for nm, base, size in regions:
                if base <= v < base + size:
                    return f"={nm}+{v - base}"
            if v not in names:
                names[v] = f"t{len(names)}"
            return "=" + names[v]

        ln = PTR.sub(sub, ln)
        m = re.search(r"\bev=(\d+)", ln)
        if m:
            ev_names.setdefault(m.group(1), f"e{len(ev_names)}")
            ln = ln.replace(f"ev={m.group(1)}", f"ev={ev_names[m.group(1)]}")
        out.append(ln)
    return out


def _pair(n, d):
    g = torch.Generator().manual_seed(n * 7 + d)
    A = torch.randn(n, d, generator=g).to(torch.bfloat16)
    B = torch.randn(n, d, generator=g).to(torch.bfloat16)
    return A, B, torch.ones(1, dtype=torch.float32)


def _sym_regions(comm):
    total = comm.buf.numel()
    return [("sym", comm.buf.data_ptr(), total), ("mc", MC_BASE, total)]


def _assert_same(py, cs):
    assert len(py) == len(cs), "\n".join(["python path:"] + py + ["sequencer:"] + cs)
    for i, (x, y) in enumerate(zip(py, cs)):
        assert x == y, f"line {i}:\n  python   : {x}\n  sequencer: {y}"


# ---------------------------------------------------------------------------------------------
# world 1
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,panel_rows", [(512, 64, None), (300, 72, None), (640, 128, 256), (1000, 64, 384)])
@pytest.mark.parametrize("need", [(True, True), (True, False), (False, True)])
def test_single_gpu_sequence_equals_python_path(streams, n, d, panel_rows, need):
    A, B, scale = _pair(n, d)
    ldw = (n + 63) // 64 * 64
    pb = cl.DEFAULT_PANEL_BYTES if panel_rows is None else 2 * ldw * panel_rows
    py, reg = _run(streams, A, B, scale, _cfg(1, 0, False, False, pb, False), comm_mod.LocalComm(K), need)
    cs, reg2 = _run(streams, A, B, scale, _cfg(1, 0, False, False, pb, True), comm_mod.LocalComm(K), need)
    cpy, ccs = _canon(py, reg), _canon(cs, reg2)
    _assert_same(cpy, ccs)
    kinds = [ln.split()[0] for ln in ccs]
    n_panels = 1 if panel_rows is None else -(-n // panel_rows)
    assert kinds.count("dz_panel") == n_panels
    assert kinds.count("gemm") == n_panels * (int(need[0]) + int(need[1]))
    assert kinds[:3] == ["rowstats", "fwd_sums", "loss_finalize"]
    # what torch.zeros does on the Python path: header (loss, maxima, flag), finalize scratch, sums
    ms = [ln for ln in cs if ln.startswith("memset")]
    assert [int(re.search(r"bytes=(\d+)", ln).group(1)) for ln in ms] == [64, 512, 3 * n * 4]


@pytest.mark.parametrize("n,d", [(512, 64), (300, 72), (1000, 64)])
@pytest.mark.parametrize("need", [(True, True), (True, False), (False, True)])
def test_single_gpu_kept_exponentials_sequence_equals_python_path(streams, n, d, need):
    """keep_exp: forward keeps E, backward = one in-place rescale + one GEMM per gradient, whatever panel_bytes says."""
    A, B, scale = _pair(n, d)
    pb = 2 * ((n + 63) // 64 * 64) * 256
    py, reg = _run(streams, A, B, scale, _cfg(1, 0, False, False, pb, False, True), comm_mod.LocalComm(K), need)
    cs, reg2 = _run(streams, A, B, scale, _cfg(1, 0, False, False, pb, True, True), comm_mod.LocalComm(K), need)
    cpy, ccs = _canon(py, reg), _canon(cs, reg2)
    _assert_same(cpy, ccs)
    kinds = [ln.split()[0] for ln in ccs]
    assert kinds.count("dz_from_exp") == 1 and kinds.count("dz_panel") == 0 and kinds.count("keep") == 1
    assert kinds.count("gemm") == int(need[0]) + int(need[1])
    # the panel the forward kept is the one the backward rescales and both GEMMs read
    e = re.search(r"keep E=(\S+)", next(ln for ln in ccs if ln.split()[0] == "keep")).group(1)
    assert f"E={e} " in next(ln for ln in ccs if ln.startswith("dz_from_exp"))
    assert all(f"A={e} " in ln for ln in ccs if ln.startswith("gemm"))


@pytest.mark.parametrize("world,rank", [(2, 1), (8, 5)])
@pytest.mark.parametrize("local_loss,gwg", [(False, True), (False, False), (True, True)])
@pytest.mark.parametrize("need", [(True, True), (False, True)])
def test_nvls_kept_exponentials_sequence_equals_python_path(streams, world, rank, local_loss, gwg, need):
    n, d = 512, 64
    A, B, scale = _pair(n, d)
    traces = []
    for seq in (False, True):
        comm = FakeNvlsComm(world, rank, streams)
        lines, reg = _run(streams, A, B, scale, _cfg(world, rank, local_loss, gwg, 2 * world * n * 128, seq, True), comm, need)
        traces.append((lines, reg + _sym_regions(comm)))
    (py, rpy), (cs, rcs) = traces
    cpy, ccs = _canon(py, rpy), _canon(cs, rcs)
    _assert_same(cpy, ccs)
    kinds = [ln.split()[0] for ln in ccs]
    assert kinds.count("dz_from_exp") == 1 and kinds.count("dz_panel") == 0 and kinds.count("gemm") == 1 + int(need[0])


def test_sequencer_is_not_used_outside_its_scope(streams):
    A, B, scale = _pair(256, 64)
    s = scale.clone().requires_grad_(True)            # d logit_scale: Python path
    lines, _ = _run(streams, A, B, s, _cfg(1, 0, False, False, cl.DEFAULT_PANEL_BYTES, True), comm_mod.LocalComm(K))
    assert not any(ln.startswith("memset") for ln in lines) and any(ln.startswith("sum_f32") for ln in lines)
    A32 = A.float()                                   # fp32 features: limb-split path
    lines, _ = _run(streams, A32, B.float(), scale, _cfg(1, 0, False, False, cl.DEFAULT_PANEL_BYTES, True), comm_mod.LocalComm(K))
    assert any(ln.startswith("split_fp32") for ln in lines) and not any(ln.startswith("memset") for ln in lines)


# ---------------------------------------------------------------------------------------------
# NVLS provider, world 2 and 8 (one rank's trace; the exchanges are only recorded)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world,rank", [(2, 0), (2, 1), (8, 5)])
@pytest.mark.parametrize("local_loss,gwg", [(False, True), (False, False), (True, True)])
@pytest.mark.parametrize("n,panel_rows,need", [(512, None, (True, True)), (1024, 384, (True, True)), (1024, 384, (False, True)),
                                               (512, None, (True, False))])
def test_nvls_sequence_equals_python_path(streams, world, rank, local_loss, gwg, n, panel_rows, need):
    d = 64
    A, B, scale = _pair(n, d)
    N = world * n
    ldw = (N + 63) // 64 * 64
    pb = cl.DEFAULT_PANEL_BYTES if panel_rows is None else 2 * ldw * panel_rows
    traces = []
    for seq in (False, True):
        comm = FakeNvlsComm(world, rank, streams)
        lines, reg = _run(streams, A, B, scale, _cfg(world, rank, local_loss, gwg, pb, seq), comm, need)
        traces.append((lines, reg + _sym_regions(comm), comm))
    (py, rpy, _), (cs, rcs, comm) = traces
    cpy, ccs = _canon(py, rpy), _canon(cs, rcs)
    _assert_same(cpy, ccs)
    kinds = [ln.split()[0] for ln in ccs]
    assert kinds.count("barrier") == 4            # allocation, sums, upstream gradients, partial dB
    assert kinds.count("mc_allreduce_f32") == 2 and kinds.count("mc_reduce_bf16") == 1
    # the exchange of the upstream gradients overlaps the dL/dZ kernel only in the global modes
    g_exchange = next(ln for ln in ccs if ln.startswith("mc_allreduce_f32") and "count=%d " % ((world + 3) // 4 * 4) in ln)
    assert g_exchange.endswith("st=side" if not local_loss else "st=main")
    # memsets / copies of the C path = the torch ops of the Python path
    sym = comm.buf.data_ptr()
    small0 = sym + comm._off_small(0)
    ms = [(int(re.search(r"p=(0x[0-9a-f]+)", ln).group(1), 16), int(re.search(r"bytes=(\d+)", ln).group(1)))
          for ln in cs if ln.startswith("memset")]
    assert ms[1][1] == 512 and ms[0][1] == 64
    assert ms[2] == (small0, (comm.SUMS_AT + 3 * N) * 4)                       # small.zero_()
    W4 = (world + 3) // 4 * 4
    assert ms[-1] == (small0 + comm.G_AT * 4, W4 * 4)                          # slot.zero_()
    cp = [ln for ln in cs if ln.startswith("copy")]
    assert len(cp) == 1 and f"dst={small0 + comm.G_AT * 4 + 4 * rank:#x}" in cp[0] and "bytes=4" in cp[0]


def test_panel_split_matches_python_host():
    """oneprot_seq_bwd_panels mirrors the panel balancing of clip_loss.py::_backward_impl."""
    import ctypes
    from oneprot_b200 import _lib
    lib = _lib.load()
    for n, N, d, pb in [(32768, 32768, 1024, 5 << 28), (32768, 32768, 1024, 1 << 30), (4096, 32768, 1024, 5 << 28),
                        (1000, 1000, 64, 2 * 1024 * 384), (8192, 65536, 1024, 5 << 28), (25, 25, 64, 5 << 28),
                        (16384, 16384, 256, 1 << 28)]:
        ldw = (N + 63) // 64 * 64
        rows_cap = max(128, (pb // (2 * ldw)) // 128 * 128)
        if rows_cap < n:
            unit = K.panel_row_unit(d)
            n_panels = -(-n // rows_cap)
            target = -(-n // n_panels)
            if unit <= rows_cap:
                up = -(-target // unit) * unit
                rows_cap = up if up <= rows_cap else rows_cap // unit * unit
            else:
                rows_cap = min(rows_cap, -(-target // 128) * 128)
        panels = [(r0, min(rows_cap, n - r0)) for r0 in range(0, n, rows_cap)]
        wz_rows = min(rows_cap, (n + 127) // 128 * 128)
        rp, wz = ctypes.c_int(), ctypes.c_int()
        cnt = lib.oneprot_seq_bwd_panels(n, N, d, pb, ctypes.byref(rp), ctypes.byref(wz))
        assert cnt == len(panels) and wz.value == wz_rows
        assert [(r0, min(rp.value, n - r0)) for r0 in range(0, n, rp.value)] == panels


def test_sequencer_argument_validation():
    import ctypes as C
    from oneprot_b200 import _lib
    lib = _lib.load()
    f = _lib.FwdSeq()
    assert lib.oneprot_seq_fwd(C.byref(f)) == 1 and b"seq_fwd_begin" in lib.oneprot_last_error()
    q = _lib.BwdSeq()
    assert lib.oneprot_seq_bwd_main(C.byref(q)) == 1
    assert lib.oneprot_seq_fwd_ws_bytes(0, 0) == 0
    assert lib.oneprot_seq_fwd_ws_bytes(4096, 32768) >= lib.oneprot_clip_fwd_scratch_bytes(4096, 32768) + 3 * 32768 * 4


def _args(line):
    return {k: v for k, v in (tok.split("=", 1) for tok in line.split()[1:] if "=" in tok)}


def _addr(v):
    return 0 if v == "(nil)" else int(v, 16)


@pytest.mark.parametrize("n,d,panel_rows", [(512, 64, None), (1000, 72, 384), (640, 128, 256)])
def test_sequencer_workspace_regions_do_not_overlap(streams, n, d, panel_rows):
    """The launch-trace equality above cannot see whether two slices of the C side's single workspace
    overlap; here the extents of every slice are rebuilt from the traced addresses and checked."""
    from oneprot_b200 import _lib
    lib = _lib.load()
    A, B, scale = _pair(n, d)
    N = n
    ldw = (N + 63) // 64 * 64
    pb = cl.DEFAULT_PANEL_BYTES if panel_rows is None else 2 * ldw * panel_rows
    lines, _ = _run(streams, A, B, scale, _cfg(1, 0, False, False, pb, True), comm_mod.LocalComm(K))
    cut = lines.index("---- backward")
    fwd, bwd = lines[:cut], lines[cut + 1:]
    # ---- forward: [finalize scratch 512 | sums 3N floats (padded) | forward scratch]
    ms = [_args(ln) for ln in fwd if ln.startswith("memset")]
    fs = _args(next(ln for ln in fwd if ln.startswith("fwd_sums")))
    fin = _args(next(ln for ln in fwd if ln.startswith("loss_finalize")))
    ws0 = _addr(ms[1]["p"])                                       # second memset = finalize scratch = workspace base
    regions = [("fin", _addr(fin["scratch"]), 512), ("sums", _addr(fs["colsum"]), 3 * N * 4),
               ("fscratch", _addr(fs["scratch"]), int(lib.oneprot_clip_fwd_scratch_bytes(n, N)))]
    assert _addr(fs["rowsum"]) == _addr(fs["colsum"]) + 4 * N and _addr(fin["diag"]) == _addr(fs["colsum"]) + 8 * N
    end = ws0 + int(lib.oneprot_seq_fwd_ws_bytes(n, N))
    regions.sort(key=lambda r: r[1])
    assert regions[0][1] >= ws0
    for (na, a0, sz), (nb, b0, _) in zip(regions, regions[1:]):
        assert a0 + sz <= b0, (na, nb)
    assert regions[-1][1] + regions[-1][2] <= end
    saved = _addr(fin["loss"])
    assert _addr(fin["stats"]) == saved + 16 and _addr(fin["flag"]) == saved + 32
    assert _addr(fin["inv_rs"]) == saved + 64 and _addr(fin["inv_cs"]) == saved + 64 + 4 * N
    # ---- backward: [g placeholder 256 | g gathered 256 | wr dg sA wc sB | fp32 dB accumulator | Wz panel]
    bw = _args(next(ln for ln in bwd if ln.startswith("bwd_weights")))
    dz = [_args(ln) for ln in bwd if ln.startswith("dz_panel")]
    gm = [_args(ln) for ln in bwd if ln.startswith("gemm")]
    n_panels = len(dz)
    rows_cap = max(int(a["rows"]) for a in dz)
    regs = [("wr", _addr(bw["wr"]), 4 * n), ("dg", _addr(bw["dg"]), 4 * n), ("sA", _addr(bw["sA"]), 4 * n),
            ("wc", _addr(bw["wc"]), 4 * N), ("sB", _addr(bw["sB"]), 4 * N), ("Wz", _addr(dz[0]["Wz"]), rows_cap * ldw * 2)]
    accs = {_addr(a["acc_out"]) for a in gm if _addr(a["acc_out"])}
    assert len(accs) == (1 if n_panels > 1 else 0)
    if accs:
        regs.append(("acc", accs.pop(), N * d * 4))
    ws_b = min(r[1] for r in regs) - 512
    endb = ws_b + int(lib.oneprot_seq_bwd_ws_bytes(n, N, d, 1, 1, pb))
    regs.sort(key=lambda r: r[1])
    for (na, a0, sz), (nb, b0, _) in zip(regs, regs[1:]):
        assert a0 + sz <= b0, (na, nb)
    assert regs[-1][1] + regs[-1][2] <= endb
    assert all(_addr(a["Wz"]) == _addr(dz[0]["Wz"]) for a in dz) and all(int(a["ldw"]) == ldw for a in dz)


# ---------------------------------------------------------------------------------------------
# static NVLS provider (CUDA-graph replay with world_size > 1): a step must be the SAME command list at every call
# ---------------------------------------------------------------------------------------------
class FakeStaticNvlsComm(FakeNvlsComm):
    def __init__(self, world, rank, streams):
        comm_mod.NvlsComm.__init__(self, K, object(), world, rank, torch.device("cpu"), static=True)
        self.symm = FakeSymm(world, streams)
        self.streams = streams

    def side_stream(self, dev):
        return comm_mod.NvlsComm.side_stream(self, dev)      # static: None (no side stream inside a captured step)


@pytest.mark.parametrize("ll,gwg", [(False, True), (True, True), (False, False)])
def test_static_provider_repeats_one_command_list(streams, ll, gwg):
    """What graphed.py captures at world_size > 1: every step starts with a barrier (write-after-read protection of the
    single buffer set), gathers with the separate multicast kernel (no epoch-flagged fused gather), stays on one stream,
    runs 5 barriers, and - addresses included - is identical from one call to the next."""
    world, n, d = 2, 512, 64                      # n % 256 == 0: the eager provider would fuse the gather here
    A, B, scale = _pair(n, d)
    comm = FakeStaticNvlsComm(world, 0, streams)
    cfg = dict(_cfg(world, 0, ll, gwg, cl.DEFAULT_PANEL_BYTES, False, True))
    runs = []
    for _ in range(3):
        lines, reg = _run(streams, A, B, scale, cfg, comm)
        runs.append(_canon(lines, reg + _sym_regions(comm), drop=()))
    sym_only = [[ln for ln in r if "sym+" in ln or "mc+" in ln or ln.startswith("barrier")] for r in runs]
    assert sym_only[1] == sym_only[2]             # same symmetric-memory addresses, same order (run 0 allocates the workspace)
    ln = runs[2]
    kinds = [x.split()[0] for x in ln]
    assert kinds.count("barrier") == 5 and kinds[0] == "barrier"
    assert "mc_store" in kinds and kinds.index("mc_store") < kinds.index("fwd_sums")
    assert not any(x.startswith("  ag ") for x in ln)                           # no fused gather
    assert not any(x.endswith(f"st={SIDE:#x}") or x.endswith("st=side") for x in ln)   # one stream
