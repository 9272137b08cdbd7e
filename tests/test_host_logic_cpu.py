"""CPU tests of the host side of ClipLoss: API surface, mode conventions and the N > 1 sharding
logic (2-rank gloo).  The CUDA kernels are replaced by tests/fake_kernels.py here - this file
checks plumbing, not numerics of the product kernels (those are the -m gpu tests)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as oc
from tests import fake_kernels
from tests.helpers import bf16_from_bits, cosine, load_golden, rel_err


@pytest.fixture()
def fake(monkeypatch):
    from oneprot_b200 import clip_loss
    monkeypatch.setattr(clip_loss, "_KERNELS", fake_kernels)
    clip_loss._SCALE_CACHE.clear()
    return clip_loss


def test_module_surface_matches_reference(fake):
    m = fake.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=3, world_size=8, use_horovod=False)
    # no parameters / buffers: strict checkpoint loading of reference checkpoints keeps working (src/train.py:80)
    assert list(m.parameters()) == [] and list(m.buffers()) == [] and m.state_dict() == {}
    assert (m.local_loss, m.gather_with_grad, m.cache_labels, m.rank, m.world_size) == (True, True, True, 3, 8)
    assert m.prev_num_logits == 0 and m.labels == {}
    lab = m.get_ground_truth(torch.device("cpu"), 4)
    assert lab.tolist() == [12, 13, 14, 15] and lab.dtype == torch.long      # loss.py:76-77
    assert m.get_ground_truth(torch.device("cpu"), 4) is lab                   # cached (loss.py:78-82)
    m1 = fake.ClipLoss()
    assert m1.get_ground_truth(torch.device("cpu"), 3).tolist() == [0, 1, 2]
    with pytest.raises(ValueError):
        m1(torch.zeros(4, 8), torch.zeros(5, 8))
    with pytest.raises(ValueError):
        m1(torch.zeros(4, 8), torch.zeros(4, 8, dtype=torch.bfloat16))


def test_product_path_has_no_cpu_fallback():
    from oneprot_b200 import ClipLoss
    from oneprot_b200._lib import OneProtKernelError
    a = torch.randn(8, 16).to(torch.bfloat16)
    with pytest.raises(OneProtKernelError):
        ClipLoss()(a, a)


@pytest.mark.parametrize("name", ["clip_single_n25_d64_train.npz", "clip_single_n100_d72_scale.npz",
                                  "clip_single_n96_d128_uncorr.npz"])
@pytest.mark.parametrize("panel_bytes", [1 << 30, 128 * 128 * 2])
def test_single_rank_host_path_vs_golden(fake, name, panel_bytes):
    g = load_golden(name)
    A = bf16_from_bits(g["A_bf16"]).requires_grad_(True)
    B = bf16_from_bits(g["B_bf16"]).requires_grad_(True)
    is_t = bool(g["scale_is_tensor"])
    ls = torch.tensor(float(g["scale"]), requires_grad=True) if is_t else float(g["scale"])
    m = fake.ClipLoss(loss_dtype=torch.float32, panel_bytes=panel_bytes, keep_exp=False)
    out = m(A, B, ls, output_dict=True)
    assert set(out) == {"contrastive_loss"}
    loss = out["contrastive_loss"]
    assert loss.dim() == 0
    loss.backward()
    assert rel_err(loss.item(), g["loss_f64"]) < 1e-5
    assert cosine(A.grad.float().numpy(), g["dA_f64"]) > 0.99999
    assert cosine(B.grad.float().numpy(), g["dB_f64"]) > 0.99999
    assert A.grad.dtype == torch.bfloat16
    if is_t:
        assert rel_err(ls.grad.item(), g["dscale_f64"]) < 2e-2
    # default return dtype = input dtype (bf16 in -> bf16 out)
    assert fake.ClipLoss()(A.detach(), B.detach(), float(g["scale"])).dtype == torch.bfloat16
    assert int(m.last_hazard_flag.item()) == 0


def test_fp32_inputs_use_limb_split(fake):
    g = load_golden("clip_single_n25_d64_train.npz")
    A = bf16_from_bits(g["A_bf16"]).float()
    B = bf16_from_bits(g["B_bf16"]).float()
    A = (A + 1e-3 * torch.randn(A.shape, generator=torch.Generator().manual_seed(1))).requires_grad_(True)
    B = B.clone().requires_grad_(True)
    loss = fake.ClipLoss()(A, B)
    assert loss.dtype == torch.float32
    loss.backward()
    ref = oc.clip_loss_closed_form(A.detach().double().numpy(), B.detach().double().numpy(), 1.0)
    assert rel_err(loss.item(), ref.loss) < 1e-5          # fp32 tolerance of BASELINE.json
    assert cosine(A.grad.numpy(), ref.dA) > 0.9999 and cosine(B.grad.numpy(), ref.dB) > 0.9999
    assert A.grad.dtype == torch.float32


def test_odd_feature_dim_is_padded(fake):
    g = torch.Generator().manual_seed(2)
    A = torch.nn.functional.normalize(torch.randn(20, 13, generator=g), dim=-1).to(torch.bfloat16).requires_grad_(True)
    B = (torch.nn.functional.normalize(torch.randn(20, 13, generator=g), dim=-1) * 5).to(torch.bfloat16).requires_grad_(True)
    loss = fake.ClipLoss(loss_dtype=torch.float32)(A, B)
    loss.backward()
    ref = oc.clip_loss_closed_form(A.detach().double().numpy(), B.detach().double().numpy(), 1.0)
    assert rel_err(loss.item(), ref.loss) < 1e-5
    assert A.grad.shape == (20, 13) and cosine(A.grad.float().numpy(), ref.dA) > 0.9999


def _wild_inputs(n=96, d=32, seed=3):
    """Row maxima thousands of bits apart: far outside the single-reference window."""
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, d, generator=g)
    b = torch.randn(n, d, generator=g)
    a[: n // 2] *= 40.0
    b[: n // 3] *= 25.0
    return a.to(torch.bfloat16), b.to(torch.bfloat16)


@pytest.mark.parametrize("mode", ["always", "auto"])
def test_two_reference_path_handles_arbitrary_inputs(fake, mode):
    a, b = _wild_inputs()
    ref = oc.clip_loss_closed_form(a.double().numpy(), b.double().numpy(), 1.0)
    # the default path must refuse these inputs ...
    m0 = fake.ClipLoss(loss_dtype=torch.float32)
    m0(a, b)
    with pytest.raises(FloatingPointError):
        m0.check_last_call()
    # ... the two-reference path computes them
    A = a.clone().requires_grad_(True)
    B = b.clone().requires_grad_(True)
    ls = torch.tensor(1.0, requires_grad=True)
    m = fake.ClipLoss(loss_dtype=torch.float32, robust=mode, panel_bytes=128 * 128 * 2, keep_exp=False)
    loss = m(A, B, ls)
    loss.backward()
    m.check_last_call()
    assert rel_err(loss.item(), ref.loss) < 1e-5
    assert cosine(A.grad.float().numpy(), ref.dA) > 0.9999
    assert cosine(B.grad.float().numpy(), ref.dB) > 0.9999
    assert abs(ls.grad.item() - ref.dscale) < 2e-2 * abs(ref.dscale) + 1e-6


def test_two_reference_path_equals_default_inside_the_window(fake):
    g = load_golden("clip_single_n100_d72_scale.npz")
    outs = []
    for mode in ("off", "always", "auto"):
        A = bf16_from_bits(g["A_bf16"]).requires_grad_(True)
        B = bf16_from_bits(g["B_bf16"]).requires_grad_(True)
        loss = fake.ClipLoss(loss_dtype=torch.float32, robust=mode)(A, B, float(g["scale"]))
        loss.backward()
        outs.append((loss.item(), A.grad.float().numpy(), B.grad.float().numpy()))
        assert rel_err(loss.item(), g["loss_f64"]) < 1e-5
    assert cosine(outs[0][1], outs[1][1]) > 0.99999 and cosine(outs[0][2], outs[1][2]) > 0.99999
    assert outs[0][0] == outs[2][0]              # "auto" stays on the default path here


# ----------------------------------------------------------------------------------------------
# world_size = 2, gloo
# ----------------------------------------------------------------------------------------------
def _worker(rank, world, port, results, provider="fake"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oneprot_b200 import clip_loss
    if provider == "emu":        # the library's own kernel source under the CPU emulation (tests/emu_kernels.py)
        from tests import emu_kernels
        clip_loss._KERNELS = emu_kernels
    else:
        clip_loss._KERNELS = fake_kernels
    g = load_golden("clip_dist_w2_n12_d32.npz")
    a = bf16_from_bits(g[f"r{rank}_A_bf16"])
    b = bf16_from_bits(g[f"r{rank}_B_bf16"])
    gout = float(g["grad_outputs"][rank])
    rec = {}
    for ll in (False, True):
        for gwg in (False, True):
            for scale_grad in (False, True):
                A = a.clone().requires_grad_(True)
                B = b.clone().requires_grad_(True)
                ls = torch.tensor(float(g["scale"]), requires_grad=scale_grad)
                for robust in ("off", "always", "keep"):       # "keep": stored-exponentials backward (falls back where it must)
                    if provider == "emu" and scale_grad and robust != "off":
                        continue                                # the emulated kernels are slow: variants once per convention
                    A = a.clone().requires_grad_(True)
                    B = b.clone().requires_grad_(True)
                    ls = torch.tensor(float(g["scale"]), requires_grad=scale_grad)
                    m = clip_loss.ClipLoss(local_loss=ll, gather_with_grad=gwg, cache_labels=True, rank=rank,
                                           world_size=world, loss_dtype=torch.float32, panel_bytes=128 * 64 * 2,
                                           robust="off" if robust == "keep" else robust, keep_exp=robust == "keep")
                    loss = m(A, B, ls)
                    (loss * gout).backward()
                    tag = f"ll{int(ll)}_gwg{int(gwg)}_sg{int(scale_grad)}" + {"always": "_rob", "keep": "_keep", "off": ""}[robust]
                    rec[tag] = dict(loss=loss.item(), dA=A.grad.float().numpy(), dB=B.grad.float().numpy(),
                                    ds=(ls.grad.item() if scale_grad else None))
    # gather_features keeps the reference contract
    am, asq = clip_loss.gather_features(a.float().requires_grad_(True), b.float(), False, True, rank, world)
    rec["gather_shape"] = tuple(am.shape)
    # ranks that disagree on n must all get a ValueError before any exchange of the loss itself
    m = clip_loss.ClipLoss(rank=rank, world_size=world)
    ragged = torch.zeros(6 + rank, 32, dtype=torch.bfloat16)
    try:
        m(ragged, ragged)
        rec["ragged"] = "no error"
    except ValueError as e:
        rec["ragged"] = str(e)
    results[rank] = rec
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("provider,port", [("fake", 29733), ("emu", 29735)])
def test_two_rank_gloo_conventions_vs_reference_golden(provider, port):
    """provider "fake": float64 emulation of the kernel contracts; "emu": the real kernel source run by the
    CPU emulation - host sharding logic AND kernels (row offsets, both modes, the two-reference path)."""
    world = 2
    if provider == "emu":
        from tests import emu_kernels
        if not emu_kernels.available():
            pytest.skip("g++ or the CUDA headers are not available")
        emu_kernels.prebuild()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results, provider), nprocs=world, join=True)
    g = load_golden("clip_dist_w2_n12_d32.npz")
    for r in range(world):
        rec = results[r]
        assert rec["gather_shape"] == (24, 32)
        assert "same (n, d)" in rec["ragged"] and "(6, 32), (7, 32)" in rec["ragged"]
        for ll in (0, 1):
            for gwg in (0, 1):
                ref = f"ll{ll}_gwg{gwg}"
                for sg, rob in ((0, ""), (1, ""), (0, "_rob"), (1, "_rob"), (0, "_keep"), (1, "_keep")):
                    if provider == "emu" and sg and rob:
                        continue
                    got = rec[f"{ref}_sg{sg}{rob}"]
                    assert rel_err(got["loss"], g[f"r{r}_loss_{ref}"]) < 1e-5, (r, ref)
                    # bf16 gradients: direction AND magnitude (the W-factor conventions of SURVEY.md 8a)
                    for k in ("dA", "dB"):
                        want = g[f"r{r}_{k}_{ref}"]
                        assert cosine(got[k], want) > 0.9999, (r, ref, k)
                        assert abs(np.linalg.norm(got[k]) / np.linalg.norm(want) - 1) < 1e-2, (r, ref, k)
                    if sg:
                        assert rel_err(got["ds"], g[f"r{r}_dscale_{ref}"]) < 2e-2, (r, ref)


def test_prefetcher_rejects_cpu_device():
    from oneprot_b200.prefetch import PinnedPairPrefetcher
    with pytest.raises(ValueError):
        PinnedPairPrefetcher("cpu")


def test_retrieval_metric_matches_reference_restatement(monkeypatch):
    from oneprot_b200 import clip_loss, retrieval
    monkeypatch.setattr(clip_loss, "_KERNELS", fake_kernels)
    monkeypatch.setattr(retrieval, "_KERNELS", fake_kernels)
    g = torch.Generator().manual_seed(12)
    S = torch.nn.functional.normalize(torch.randn(300, 48, generator=g), dim=-1)
    M = torch.nn.functional.normalize(S + 0.8 * torch.randn(300, 48, generator=g), dim=-1)
    m = retrieval.RetrievalMetric(k=[1, 10, 100])
    for lo in range(0, 300, 64):                      # several update() calls, like validation batches
        m.update(S[lo:lo + 64], M[lo:lo + 64])
    got = m.compute()
    want = oc.retrieval_metric_closed_form(S.double().numpy(), M.double().numpy())
    assert set(got) == set(want)
    for k in want:
        assert abs(float(got[k]) - float(want[k])) < 1e-9, k
    m.reset()
    assert len(m.preds) == 0 and len(m.target) == 0


def test_retrieval_metric_host_logic_vs_reference_golden(monkeypatch):
    from oneprot_b200 import clip_loss, retrieval
    monkeypatch.setattr(clip_loss, "_KERNELS", fake_kernels)
    monkeypatch.setattr(retrieval, "_KERNELS", fake_kernels)
    g = load_golden("retrieval_metric.npz")
    for tag in ("easy", "hard"):
        S, M = bf16_from_bits(g[f"{tag}_S_bf16"]), bf16_from_bits(g[f"{tag}_M_bf16"])
        m = retrieval.RetrievalMetric()
        for lo in range(0, S.shape[0], 100):
            m.update(S[lo:lo + 100], M[lo:lo + 100])
        got = m.compute()
        for k, v in g.items():
            if k.startswith(tag + ":"):
                assert float(got[k.split(":", 1)[1]]) == float(v), (tag, k)
