"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY (see oracle/clip_oracle.py header).  Usage, from the repo root, in the
build container where /root/reference is mounted:

    python oracle/make_golden.py

It imports ``src.models.components.loss`` (ClipLoss, gather_features) and
``src.models.components.base_encoder`` (Normalize, LearnableLogitScaling) from /root/reference,
runs them in float64 / float32 / bfloat16 on bf16-valued synthetic inputs (single process and
2-rank gloo) and stores inputs + outputs.  Nothing at test or bench time reads /root/reference;
the fixtures are the pin for oracle/clip_oracle.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REF = os.environ.get("ONEPROT_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _ref():
    sys.path.insert(0, REF)
    from src.models.components.loss import ClipLoss, gather_features  # noqa: F401
    from src.models.components.base_encoder import Normalize, LearnableLogitScaling  # noqa: F401
    return ClipLoss, gather_features, Normalize, LearnableLogitScaling


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


def single_process_cases():
    from oracle.clip_oracle import synthetic_pair
    ClipLoss, _, _, _ = _ref()
    out = {}
    cases = [
        # name, n, d, correlated, temperature_into_b, logit_scale (None => python float 1.0)
        ("n25_d64_train", 25, 64, True, True, None),
        ("n96_d128_uncorr", 96, 128, False, True, None),
        ("n100_d72_scale", 100, 72, True, False, 1.0 / 0.07),
        ("n256_d512_train", 256, 512, True, True, None),          # BASELINE configs[0] shape
        ("n256_d512_scale", 256, 512, True, False, 1.0 / 0.07),
    ]
    for name, n, d, corr, t_in_b, s in cases:
        a, b = synthetic_pair(n, d, seed=1234, pair_id=0, rank=0, correlated=corr,
                              temperature_into_b=t_in_b, dtype="bf16")
        rec = {"A_bf16": bf16_bits(a), "B_bf16": bf16_bits(b),
               "scale": np.float64(1.0 if s is None else s), "scale_is_tensor": np.bool_(s is not None)}
        for tag, td in (("f64", torch.float64), ("f32", torch.float32), ("bf16", torch.bfloat16)):
            A = a.to(td).requires_grad_(True)
            B = b.to(td).requires_grad_(True)
            if s is None:
                ls = 1.0
            else:
                ls = torch.tensor(s, dtype=td, requires_grad=True)
            loss = ClipLoss(world_size=1)(A, B, ls)
            loss.backward()
            rec[f"loss_{tag}"] = np.float64(loss.detach().double().item())
            if tag == "f64":
                # the second 256x512 case keeps only the first 8 gradient rows (fixture size)
                keep = 8 if name == "n256_d512_scale" else n
                rec["dA_f64"] = A.grad.numpy()[:keep].astype(np.float32)
                rec["dB_f64"] = B.grad.numpy()[:keep].astype(np.float32)
                rec["dscale_f64"] = np.float64(0.0 if s is None else ls.grad.item())
            if tag == "bf16":
                rec["loss_bf16_dtype"] = np.bytes_(str(loss.dtype))
        out[name] = rec
    for name, rec in out.items():
        np.savez_compressed(os.path.join(OUT, f"clip_single_{name}.npz"), **rec)
        print("wrote", name, {k: v for k, v in rec.items() if np.ndim(v) == 0})


def _dist_worker(rank, world, port, n, d, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    ClipLoss, _, _, _ = _ref()
    from oracle.clip_oracle import synthetic_pair
    a, b = synthetic_pair(n, d, seed=4321, pair_id=0, rank=rank, correlated=True,
                          temperature_into_b=False, dtype="bf16")
    rec = {"A_bf16": bf16_bits(a), "B_bf16": bf16_bits(b)}
    for local_loss in (False, True):
        for gwg in (False, True):
            A = a.double().requires_grad_(True)
            B = b.double().requires_grad_(True)
            ls = torch.tensor(1.0 / 0.07, dtype=torch.float64, requires_grad=True)
            loss = ClipLoss(local_loss=local_loss, gather_with_grad=gwg, cache_labels=True,
                            rank=rank, world_size=world)(A, B, ls)
            # distinct upstream gradient per rank exercises the reduce-scatter convention
            (loss * (1.0 + 0.5 * rank)).backward()
            tag = f"ll{int(local_loss)}_gwg{int(gwg)}"
            rec[f"loss_{tag}"] = np.float64(loss.item())
            rec[f"dA_{tag}"] = A.grad.numpy().copy()
            rec[f"dB_{tag}"] = B.grad.numpy().copy()
            rec[f"dscale_{tag}"] = np.float64(ls.grad.item())
    results[rank] = rec
    dist.barrier()
    dist.destroy_process_group()


def distributed_cases():
    world, n, d = 2, 12, 32
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_dist_worker, args=(world, 29611, n, d, results), nprocs=world, join=True)
    flat = {"world": np.int64(world), "n": np.int64(n), "d": np.int64(d),
            "scale": np.float64(1.0 / 0.07), "grad_outputs": np.array([1.0, 1.5])}
    for r in range(world):
        for k, v in results[r].items():
            flat[f"r{r}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "clip_dist_w2_n12_d32.npz"), **flat)
    print("wrote distributed", {k: float(v) for k, v in flat.items() if k.startswith("r") and "loss" in k})


def epilogue_cases():
    _, _, Normalize, LearnableLogitScaling = _ref()
    g = torch.Generator().manual_seed(77)
    x = (3.0 * torch.randn(8, 64, generator=g)).to(torch.bfloat16)
    x[3] = 0  # exercises the eps clamp of F.normalize
    gy = torch.randn(8, 64, generator=g).to(torch.bfloat16)
    X = x.double().requires_grad_(True)
    y = Normalize(dim=-1)(X)
    y.backward(gy.double())
    scl = LearnableLogitScaling(logit_scale_init=1 / 0.07, learnable=True)
    ys = scl(y.detach().float())
    big = LearnableLogitScaling(logit_scale_init=250.0, learnable=False)   # clipped to 100
    yb = big(y.detach().float())
    np.savez_compressed(os.path.join(OUT, "epilogue_normalize_scale.npz"),
                        x_bf16=bf16_bits(x), gy_bf16=bf16_bits(gy),
                        y_f64=y.detach().numpy(), gx_f64=X.grad.numpy(),
                        log_logit_scale=np.float64(scl.log_logit_scale.item()),
                        ys_f32=ys.detach().numpy(),
                        log_logit_scale_big=np.float64(big.log_logit_scale.item()),
                        yb_f32=yb.detach().numpy())
    print("wrote epilogue")


def _siglip_worker(rank, world, port, n, d, bidir, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sys.path.insert(0, REF)
    from src.models.components.loss import SigLipLoss
    from oracle.clip_oracle import synthetic_pair
    a, b = synthetic_pair(n, d, seed=777, pair_id=0, rank=rank, correlated=True, temperature_into_b=False, dtype="bf16")
    A = a.double().requires_grad_(True)
    B = b.double().requires_grad_(True)
    st = torch.tensor(10.0, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(-10.0, dtype=torch.float64, requires_grad=True)
    loss = SigLipLoss(rank=rank, world_size=world, bidir=bidir)(A, B, st, bt)
    (loss * (1.0 + 0.5 * rank)).backward()
    results[rank] = {"A_bf16": bf16_bits(a), "B_bf16": bf16_bits(b), "loss": np.float64(loss.item()),
                     "dA": A.grad.numpy().copy(), "dB": B.grad.numpy().copy(),
                     "dscale": np.float64(st.grad.item()), "dbias": np.float64(bt.grad.item())}
    dist.barrier()
    dist.destroy_process_group()


def siglip_cases():
    """SigLipLoss of the unmodified reference (loss.py:204-311): single process (with / without
    logit_bias) and the neighbour-exchange rings on 2 and 3 gloo ranks, uni- and bidirectional, with a
    distinct upstream gradient per rank."""
    sys.path.insert(0, REF)
    from src.models.components.loss import SigLipLoss
    from oracle.clip_oracle import synthetic_pair
    rec = {}
    for tag, n, d, scale, bias in (("train", 40, 64, 1.0, None), ("paper", 96, 128, 10.0, -10.0), ("odd", 25, 72, 3.0, 0.5)):
        a, b = synthetic_pair(n, d, seed=2468, pair_id=0, rank=0, correlated=True, temperature_into_b=(tag == "train"), dtype="bf16")
        A = a.double().requires_grad_(True)
        B = b.double().requires_grad_(True)
        st = torch.tensor(float(scale), dtype=torch.float64, requires_grad=True)
        bt = None if bias is None else torch.tensor(float(bias), dtype=torch.float64, requires_grad=True)
        loss = SigLipLoss(world_size=1)(A, B, st, bt)
        loss.backward()
        rec[f"{tag}_dscale"] = np.float64(st.grad.item())
        rec[f"{tag}_dbias"] = np.float64(0.0 if bt is None else bt.grad.item())
        rec.update({f"{tag}_A_bf16": bf16_bits(a), f"{tag}_B_bf16": bf16_bits(b), f"{tag}_scale": np.float64(scale),
                    f"{tag}_bias": np.float64(0.0 if bias is None else bias), f"{tag}_has_bias": np.bool_(bias is not None),
                    f"{tag}_loss": np.float64(loss.item()), f"{tag}_dA": A.grad.numpy().copy(), f"{tag}_dB": B.grad.numpy().copy()})
        lb = SigLipLoss(world_size=1)(a.clone(), b.clone(), scale, bias)
        rec[f"{tag}_loss_bf16"] = np.float64(lb.double().item())
        rec[f"{tag}_loss_bf16_dtype"] = np.bytes_(str(lb.dtype))
    np.savez_compressed(os.path.join(OUT, "siglip_single.npz"), **rec)
    print("wrote siglip single", {k: float(v) for k, v in rec.items() if k.endswith("_loss")})
    for world, bidir, port in ((2, True, 29631), (3, True, 29633), (3, False, 29635)):
        n, d = 12, 32
        mgr = mp.Manager()
        results = mgr.dict()
        mp.spawn(_siglip_worker, args=(world, port, n, d, bidir, results), nprocs=world, join=True)
        flat = {"world": np.int64(world), "n": np.int64(n), "d": np.int64(d), "scale": np.float64(10.0), "bias": np.float64(-10.0),
                "grad_outputs": np.array([1.0 + 0.5 * r for r in range(world)])}
        for r in range(world):
            for k, v in results[r].items():
                flat[f"r{r}_{k}"] = v
        np.savez_compressed(os.path.join(OUT, f"siglip_dist_w{world}_{'bidir' if bidir else 'ring'}.npz"), **flat)
        print("wrote siglip dist", world, bidir, [float(flat[f"r{r}_loss"]) for r in range(world)])


def _stub_torchmetrics():
    """Minimal stand-ins for the parts of torchmetrics the reference imports (not installed here): list
    states for Metric, running mean / min for MeanMetric / MinMetric.  No arithmetic of the path."""
    import types
    if "torchmetrics" in sys.modules:
        return

    class Metric:
        def __init__(self, **kwargs):
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default
            setattr(self, name, list(default))

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, list(v))

    class MeanMetric:
        def __init__(self):
            self.reset()

        def reset(self):
            self.total, self.count = 0.0, 0

        def __call__(self, v):
            self.total += float(v)
            self.count += 1

        def compute(self):
            return torch.tensor(self.total / max(self.count, 1))

    class MinMetric(MeanMetric):
        def reset(self):
            self.best = float("inf")

        def __call__(self, v):
            self.best = min(self.best, float(v))

        def compute(self):
            return torch.tensor(self.best)

    tm = types.ModuleType("torchmetrics")
    tm.MeanMetric, tm.MinMetric = MeanMetric, MinMetric
    mods = {"torchmetrics": tm, "torchmetrics.metric": types.ModuleType("torchmetrics.metric"),
            "torchmetrics.utilities": types.ModuleType("torchmetrics.utilities"),
            "torchmetrics.utilities.data": types.ModuleType("torchmetrics.utilities.data"),
            "torchmetrics.utilities.imports": types.ModuleType("torchmetrics.utilities.imports"),
            "torchmetrics.utilities.plot": types.ModuleType("torchmetrics.utilities.plot")}
    mods["torchmetrics.metric"].Metric = Metric
    mods["torchmetrics.utilities"].rank_zero_warn = lambda *a, **k: None
    mods["torchmetrics.utilities.data"].dim_zero_cat = lambda x: torch.cat(list(x), dim=0) if isinstance(x, (list, tuple)) else x
    mods["torchmetrics.utilities.imports"]._MATPLOTLIB_AVAILABLE = False
    mods["torchmetrics.utilities.plot"]._AX_TYPE = object
    mods["torchmetrics.utilities.plot"]._PLOT_OUT_TYPE = object
    sys.modules.update(mods)


def _stub_lightning():
    """pytorch_lightning is not installed here: a LightningModule stand-in with exactly the plumbing
    OneProtLitModule uses in manual-optimisation mode (oneprot_module.py:22-24,82,105-108):
    save_hyperparameters, optimizers(), manual_backward, clip_gradients (norm), log, global_step."""
    import inspect
    import types
    if "pytorch_lightning" in sys.modules:
        return

    class LightningModule(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._opt, self.global_step, self.logged = None, 0, {}

        def save_hyperparameters(self, logger=False):
            loc = inspect.currentframe().f_back.f_locals
            self.hparams = types.SimpleNamespace(**{k: v for k, v in loc.items() if k not in ("self", "__class__")})

        def optimizers(self):
            if self._opt is None:
                cfg = self.configure_optimizers()
                self._opt = cfg["optimizer"] if isinstance(cfg, dict) else cfg
            return self._opt

        def manual_backward(self, loss):
            loss.backward()

        def clip_gradients(self, opt, gradient_clip_val=None, gradient_clip_algorithm=None):
            assert gradient_clip_algorithm == "norm"
            params = [p for grp in opt.param_groups for p in grp["params"]]
            torch.nn.utils.clip_grad_norm_(params, gradient_clip_val)

        def log(self, name, value, **kw):
            self.logged[name] = value
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl


def module_cases():
    """The reference's OWN OneProtLitModule (oneprot_module.py) - training_step with the L1 term,
    validation_step with RetrievalMetric, test_step with the tensor logit_scale - over the reference's
    BaseEncoder heads, in float64, world_size 1.  Lightning / torchmetrics plumbing is stubbed (above);
    ClipLoss, BaseEncoder, RetrievalMetric and the step logic are the unmodified reference."""
    import functools
    _stub_torchmetrics()
    _stub_lightning()
    os.environ["RANK"], os.environ["WORLD_SIZE"] = "0", "1"
    sys.path.insert(0, REF)
    from src.models.components.base_encoder import BaseEncoder
    from src.models.oneprot_module import OneProtLitModule
    g = torch.Generator().manual_seed(4242)
    spec = {"sequence": (64, "mlp", False, "mean"), "text": (48, "mlp", True, "cls"), "struct_graph": (56, "linear", True, "mean")}
    comps = {k: BaseEncoder(dm, 32, proj_type=pt, use_logit_scale=uls, learnable_logit_scale=False, pooling_type=pool).double()
             for k, (dm, pt, uls, pool) in spec.items()}
    with torch.no_grad():
        for enc in comps.values():
            for p in enc.parameters():
                p.copy_((p + 0.2 * torch.randn(p.shape, generator=g).double()).to(torch.bfloat16).double())
    module = OneProtLitModule(comps, optimizer=functools.partial(torch.optim.SGD, lr=0.02, momentum=0.9), loss_fn="CLIP",
                              use_l1_regularization=True, local_loss=True, gather_with_grad=True)
    rec = {"spec": np.bytes_(repr(spec)), "lr": np.float64(0.02), "momentum": np.float64(0.9)}
    for k, v in module.network.state_dict().items():
        rec["init:" + k] = v.numpy().copy()
    B, L, steps = 12, 5, 4

    def make_batch(tag):
        b = {}
        for mod in ("text", "struct_graph"):
            seq = torch.randn(B, L, 64, generator=g).to(torch.bfloat16)
            x = torch.randn((B, L, 48) if mod == "text" else (B, 56), generator=g).to(torch.bfloat16)
            x = x + 0.5 * seq.float().mean(1)[:, :x.shape[-1]].reshape((B, 1, -1) if x.dim() == 3 else (B, -1)).to(torch.bfloat16)
            rec[f"{tag}:{mod}:seq_bf16"], rec[f"{tag}:{mod}:mod_bf16"] = bf16_bits(seq), bf16_bits(x)
            rec[f"{tag}:{mod}:seq_shape"], rec[f"{tag}:{mod}:mod_shape"] = np.array(seq.shape), np.array(x.shape)
            b[mod] = (seq.double(), x.double(), None, None)
        return b

    # the reference's training_step only logs the running mean: capture each call's loss value
    losses = []
    orig = module.train_loss.__call__
    module.train_loss = type("Tap", (), {"__call__": lambda self, v: losses.append(float(v)), "reset": lambda self: None})()
    for s in range(steps):
        module.training_step(make_batch(f"train{s}"))
        module.global_step += 1
    rec["train_losses"] = np.array(losses)                       # order: step-major, (text, struct_graph)
    for k, v in module.network.state_dict().items():
        rec["final:" + k] = v.numpy().copy()
    vb = make_batch("val")
    vals = []
    module.val_loss = type("Tap", (), {"__call__": lambda self, v: vals.append(float(v)), "reset": lambda self: None})()
    with torch.no_grad():
        for mod in ("text", "struct_graph"):
            seq, x, _, _ = vb[mod]
            module.validation_step((seq, x, mod, None), 0)
    rec["val_losses"] = np.array(vals)
    for mod in ("text", "struct_graph"):
        for k, v in module.metrics["val_" + mod].compute().items():
            rec[f"valmetric:{mod}:{k}"] = np.float64(v)
    tests = []
    module.test_loss = type("Tap", (), {"__call__": lambda self, v: tests.append(float(v)), "reset": lambda self: None})()
    with torch.no_grad():
        module.test_step(vb, 0)
    rec["test_losses"] = np.array(tests)
    np.savez_compressed(os.path.join(OUT, "module_steps.npz"), **rec)
    print("wrote module", losses, vals, tests)


def retrieval_cases():
    """RetrievalMetric.update / compute of the unmodified reference (retrieval_metric.py:71-102).
    torchmetrics is not installed in the build container, so its ``Metric`` base class and helpers are
    replaced by minimal stand-ins that only keep the list states (``add_state`` / ``dim_zero_cat``) -
    the arithmetic under test (similarity, argsort, rank of the label, median, R@k) is the reference's."""
    _stub_torchmetrics()
    sys.path.insert(0, REF)
    from src.models.components.retrieval_metric import RetrievalMetric
    rec = {}
    g = torch.Generator().manual_seed(99)
    for tag, n, d, noise in (("easy", 300, 64, 0.5), ("hard", 257, 32, 2.5)):
        S = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).to(torch.bfloat16)
        M = torch.nn.functional.normalize(S.float() + noise * torch.randn(n, d, generator=g), dim=-1).to(torch.bfloat16)
        m = RetrievalMetric()
        for lo in range(0, n, 100):                      # several update() calls, as the validation loop does
            m.update(S[lo:lo + 100].double(), M[lo:lo + 100].double())
        out = m.compute()
        rec[f"{tag}_S_bf16"], rec[f"{tag}_M_bf16"] = bf16_bits(S), bf16_bits(M)
        for k, v in out.items():
            rec[f"{tag}:{k}"] = np.float64(v)
    np.savez_compressed(os.path.join(OUT, "retrieval_metric.npz"), **rec)
    print("wrote retrieval", {k: float(v) for k, v in rec.items() if ":" in k})


def head_cases():
    """BaseEncoder heads of the unmodified reference (base_encoder.py:129-194) in float64 on
    bf16-valued inputs and parameters: output + gradients w.r.t. the input and every parameter."""
    sys.path.insert(0, REF)
    from src.models.components.base_encoder import BaseEncoder
    cases = [
        # name, d_model, output_dim, proj_type, pooling, use_logit_scale, learnable, B, L
        ("mlp_mean_scale", 64, 32, "mlp", "mean", True, True, 6, 9),        # sequence / text style head
        ("linear_cls", 48, 40, "linear", "cls", False, False, 5, 7),
        ("linear_identity_scale", 56, 24, "linear", "identity", True, False, 10, 0),   # struct-graph style: 2-D input
        ("none_mean", 32, 32, None, "mean", False, False, 4, 5),
        # attention1d pooling has its hidden size hard-wired to 1280 (base_encoder.py:179); train_ddp_1.yaml's sequence head
        ("linear_attn1d", 1280, 16, "linear", "attention1d", False, False, 3, 6),
    ]
    for name, dm, do, proj, pool, uls, learn, B, L in cases:
        g = torch.Generator().manual_seed(len(name) * 101 + dm)
        enc = BaseEncoder(dm, do, proj_type=proj, use_logit_scale=uls, learnable_logit_scale=learn, pooling_type=pool).double()
        with torch.no_grad():
            for k, p in enc.state_dict().items():
                if p.dim() >= 1:      # bf16-valued parameters away from the trivial LayerNorm init
                    p.copy_((p + 0.3 * torch.randn(p.shape, generator=g).double()).to(torch.bfloat16).double())
        x = torch.randn((B, L, dm) if L else (B, dm), generator=g).to(torch.bfloat16)
        mask = None
        if L and pool in ("mean", "attention1d"):
            lens = torch.randint(1, L + 1, (B,), generator=g)
            mask = (torch.arange(L)[None, :] < lens[:, None]).long()
        gy = torch.randn(B, do, generator=g).to(torch.bfloat16)
        X = x.double().requires_grad_(True)
        # 'identity' pooling is nn.Identity, which takes no mask: the reference's encoders that use it call
        # proj / norm themselves (struct_graph_encoder.py:41-42)
        y = enc(X, mask) if pool != "identity" else enc.norm(enc.proj(X))
        y.backward(gy.double())
        rec = {"x_bf16": bf16_bits(x), "x_shape": np.array(x.shape), "gy_bf16": bf16_bits(gy),
               "mask": np.zeros(0) if mask is None else mask.numpy(), "y_f64": y.detach().numpy(), "gx_f64": X.grad.numpy(),
               "d_model": np.int64(dm), "output_dim": np.int64(do), "proj_type": np.bytes_(str(proj)),
               "pooling_type": np.bytes_(pool), "use_logit_scale": np.bool_(uls), "learnable": np.bool_(learn)}
        named = dict(enc.named_parameters())
        for k, v in enc.state_dict().items():
            rec["param:" + k] = v.detach().numpy()
            if k in named and named[k].grad is not None:
                rec["grad:" + k] = named[k].grad.numpy()
        np.savez_compressed(os.path.join(OUT, f"head_{name}.npz"), **rec)
        print("wrote head", name, sorted(k for k in rec if k.startswith("grad:")))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    only = [a for a in sys.argv[1:] if a.startswith("--") and a.endswith("-only")]
    if not only:
        single_process_cases()
        distributed_cases()
        epilogue_cases()
    if not only or "--heads-only" in only:
        head_cases()
    if not only or "--siglip-only" in only:
        siglip_cases()
    if not only or "--retrieval-only" in only:
        retrieval_cases()
    if not only or "--module-only" in only:
        module_cases()
