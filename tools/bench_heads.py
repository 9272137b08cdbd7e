"""HBM roofline of the projection-head row kernels (csrc/head_kernels.cu) and the time of a whole
BaseEncoder head fwd+bwd at OneProt's sizes (tokens B x L x 1280 -> 1152 -> 1024), CUDA events, L2
flushed between repetitions.  Peak = MEASURED_PEAKS.json hbm_gbs (fallback 6650 GB/s).

    python tools/bench_heads.py [rows] [--json gpurun_out/heads_bench.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oneprot_b200 import kernels as K  # noqa: E402
from oneprot_b200.heads import BaseEncoder  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def timed(fn, flush, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 65536
    out_path = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else os.path.join(ROOT, "gpurun_out", "heads_bench.json")
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, src = peak_gbs()
    res = {"peak_gbs": peak, "peak_source": src, "rows": rows, "kernels": {}}
    for dt, esz in ((torch.bfloat16, 2), (torch.float32, 4)):
        for d in (1280, 1152):
            x = torch.randn(rows, d, device=dev).to(dt)
            gy = torch.randn(rows, d, device=dev).to(dt)
            y = torch.empty_like(x); gx = torch.empty_like(x)
            w = torch.ones(d, device=dev, dtype=dt); b = torch.zeros(d, device=dev, dtype=dt)
            mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
            dg = torch.empty(d, device=dev); db = torch.empty(d, device=dev)
            K.layernorm_fwd(x, w, b, y, mean, rstd, 1e-5)
            cases = {
                f"layernorm_fwd d={d}": (lambda: K.layernorm_fwd(x, w, b, y, mean, rstd, 1e-5), 2 * rows * d * esz),
                f"layernorm_bwd d={d} (dx + dgamma + dbeta)": (lambda: K.layernorm_bwd(x, gy, w, mean, rstd, gx, dg, db), 3 * rows * d * esz),
                f"gelu_fwd d={d}": (lambda: K.gelu(x, y), 2 * rows * d * esz),
                f"gelu_bwd d={d}": (lambda: K.gelu(x, gx, gy), 3 * rows * d * esz),
            }
            for name, (fn, nbytes) in cases.items():
                ms = timed(fn, flush)
                gbs = nbytes / (ms * 1e-3) / 1e9
                res["kernels"][f"{name} {str(dt).split('.')[-1]}"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac": gbs / peak}
        # the epilogue in front of the loss: L2-normalise (+ logit scale) and the plain scale, d = 1024
        d = 1024
        x = torch.randn(rows, d, device=dev).to(dt)
        gy = torch.randn(rows, d, device=dev).to(dt)
        y = torch.empty_like(x); gx = torch.empty_like(x)
        inv = torch.empty(rows, device=dev); dsp = torch.empty(rows, device=dev)
        sc = torch.full((1,), 1 / 0.07, device=dev)
        K.l2norm_scale_fwd(x, y, inv, sc)
        cases = {
            f"l2norm_scale_fwd d={d}": (lambda: K.l2norm_scale_fwd(x, y, inv, sc), 2 * rows * d * esz),
            f"l2norm_scale_bwd d={d}": (lambda: K.l2norm_scale_bwd(x, gy, inv, gx, dsp, sc), 3 * rows * d * esz),
            f"scale_rows d={d}": (lambda: K.scale_rows(x, y, sc), 2 * rows * d * esz),
        }
        for name, (fn, nbytes) in cases.items():
            ms = timed(fn, flush)
            gbs = nbytes / (ms * 1e-3) / 1e9
            res["kernels"][f"{name} {str(dt).split('.')[-1]}"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac": gbs / peak}
        Bt, L, D = 64, 1024, 1280
        f = torch.randn(Bt, L, D, device=dev).to(dt)
        mask = (torch.arange(L, device=dev)[None, :] < torch.randint(L // 2, L + 1, (Bt, 1), device=dev)).float()
        yp = torch.empty(Bt, D, device=dev, dtype=dt); inv = torch.empty(Bt, device=dev)
        ms = timed(lambda: K.meanpool_fwd(f, mask, yp, inv), flush)
        nbytes = (Bt * L * D + Bt * D) * esz
        res["kernels"][f"meanpool_fwd {Bt}x{L}x{D} {str(dt).split('.')[-1]}"] = {"ms": ms, "algorithmic_bytes": nbytes,
                                                                                  "gbs": nbytes / (ms * 1e-3) / 1e9,
                                                                                  "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
    # whole head, fwd + bwd, the shipped sequence-tower head on pooled features (8192 x 1280)
    for dt in (torch.bfloat16, torch.float32):
        n = 8192
        enc = BaseEncoder(1280, 1024, proj_type="mlp", pooling_type="mean").cuda().to(dt)
        x = torch.randn(n, 1280, device=dev).to(dt).requires_grad_(True)
        gy = torch.randn(n, 1024, device=dev).to(dt)

        def step():
            x.grad = None
            for p in enc.parameters():
                p.grad = None
            enc(x).backward(gy)

        K.launch_count_reset()
        step()
        launches = K.launch_count()
        ms = timed(step, flush)
        hid = (1280 + 1024) // 2
        flops = 3 * 2.0 * n * (1280 * hid + hid * 1024)
        res[f"head_mlp_fwd_bwd n={n} {str(dt).split('.')[-1]}"] = {"ms": ms, "launches": launches, "gemm_tflops": flops / (ms * 1e-3) / 1e12}
    for k, v in res["kernels"].items():
        print(f"{k:60s} {v['ms']:8.4f} ms  {v['gbs']:8.1f} GB/s  {100 * v['frac']:5.1f} % of {peak:.0f}")
    for k, v in res.items():
        if k.startswith("head_"):
            print(k, v)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
