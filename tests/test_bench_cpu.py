"""CPU: bench.py prints exactly ONE well-formed JSON line - through the pipelined e2e leg and through its
fallback - when its GPU arm is driven by stand-ins (tests/bench_cpu_harness.py), and the reference arm
(`--impl reference`) does the same for real on the host cores."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"}


def _json_lines(out):
    return [json.loads(ln) for ln in out.splitlines() if ln.startswith("{")]


@pytest.mark.parametrize("mode", ["fallback", "pipelined"])
def test_gpu_arm_control_flow_prints_one_line(mode):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_cpu_harness.py"), mode], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = _json_lines(p.stdout)
    assert len(lines) == 1, p.stdout
    ln = lines[0]
    assert KEYS <= set(ln), KEYS - set(ln)
    assert ln["metric"] == "cliploss_fwd_bwd_samples_per_s" and ln["unit"] == "samples/s" and ln["n_gpus"] == 1
    assert ln["steps"] == 3 and ln["warmup"] == 3 and ln["scaling"] == "strong" and ln["vs_baseline"] is None
    assert "workload" in ln["config"] and ln["config"]["warmup_steps_run"] == 3
    r = ln["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "tensor"
    assert {"value", "unit", "cores", "kind", "sample"} <= set(ln["cpu_baseline"]) and ln["cpu_baseline"]["kind"] == "port"
    e = ln["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "mode", "serial_value"} <= set(e)
    assert e["h2d_bytes_per_step"] == 2 * 256 * 64 * 2 and e["d2h_bytes_per_step"] == 4
    if mode == "pipelined":      # the leg ran; it is the headline only if it beat the serial leg (CPU timing here is noise)
        assert "pipelined_ms_per_step" in e and (e["mode"].startswith("pipelined") or "was slower" in e["mode"])
    else:
        assert e["mode"].startswith("serial") and "pipelined leg failed" in e["mode"]


def test_gpu_arm_over_the_emulated_library():
    """bench.py through the product's own wrappers / C entry points (library compiled for the CPU): the launch count the
    line claims is the library's own counter, 10 kernels per step."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_cpu_harness.py"), "lib-pipelined"], capture_output=True,
                       text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = _json_lines(p.stdout)
    assert len(lines) == 1, p.stdout
    ln = lines[0]
    assert KEYS <= set(ln) and ln["gpu_launches"] == 20 and "pipelined_ms_per_step" in ln["e2e"]
    assert len(ln["roofline"]["kernels"]) == 4 and ln["roofline"]["step"]["executed_over_algorithmic"] == 1.0    # stored exponentials
    assert ln["config"]["sustained"]["steps"] >= 2 and "eager_b200" in ln
    assert 2.0 < ln["config"]["loss"] < 6.0          # N = 128 correlated pairs: below ln 128 = 4.85


def test_two_rank_control_flow_over_gloo():
    """The driver's multi-GPU launch line with 2 ranks: rank 0 prints the one line, both exit 0."""
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29671", os.path.join(ROOT, "tests", "bench_cpu_harness.py"), "pipelined"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = _json_lines(p.stdout)
    assert len(lines) == 1, p.stdout
    ln = lines[0]
    assert ln["n_gpus"] == 2 and ln["config"]["rows_per_gpu"] == 128 and ln["config"]["warmup_steps_run"] == 6 and ln["warmup"] == 3
    assert "cpu_baseline" not in ln                      # rank 0 at N = 1 only
    assert "pipelined_ms_per_step" in ln["e2e"] and ln["e2e"]["h2d_bytes_per_step"] == 2 * 128 * 64 * 2


def test_reference_arm_prints_one_line():
    env = dict(os.environ, ONEPROT_BENCH_N="512", ONEPROT_BENCH_D="64")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = _json_lines(p.stdout)
    assert len(lines) == 1
    ln = lines[0]
    assert ln["impl"] == "reference" and ln["e2e"]["h2d_bytes_per_step"] == 0 and ln["cpu_baseline"]["kind"] == "port"
    assert ln["e2e"]["value"] == ln["value"] == ln["cpu_baseline"]["value"]


def test_global_synthetic_pair_does_not_depend_on_the_sharding():
    """bench.py draws rows [rank n, (rank + 1) n) of ONE global pair: any sharding concatenates to the same matrices, so
    the global loss is comparable across GPU counts (different data per world size)."""
    import torch
    sys.path.insert(0, ROOT)
    from tools.synthetic import GLOBAL_BLOCK, synthetic_global_rows
    N, d = 4 * GLOBAL_BLOCK, 16
    a, b = synthetic_global_rows(0, N, d)
    for world in (2, 4, 8):
        n = N // world
        parts = [synthetic_global_rows(r * n, n, d) for r in range(world)]
        assert torch.equal(torch.cat([p[0] for p in parts]), a) and torch.equal(torch.cat([p[1] for p in parts]), b)
    a2, b2 = synthetic_global_rows(300, 1000, d)                       # unaligned windows are slices of the same pair
    assert torch.equal(a2, a[300:1300]) and torch.equal(b2, b[300:1300])
    a3, _ = synthetic_global_rows(0, N, d, pair_id=1)
    assert not torch.equal(a3, a)


def test_modalities5_config_control_flow():
    """`bench.py --config modalities5` (BASELINE cfg 3) through the stand-ins: one JSON line, per-host-variant times, the
    CUDA-graph variant reported as unavailable on the CPU without taking the line down."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_cpu_harness.py"), "m5"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = _json_lines(p.stdout)
    assert len(lines) == 1, p.stdout
    ln = lines[0]
    c = ln["config"]
    assert c["pairs"] == 5 and c["rows_per_gpu"] == 1024 and ln["scaling"] == "weak" and ln["n_gpus"] == 1
    assert "python host" in c["ms_per_step_by_host"] and abs(c["us_per_pair"] - 1e3 * ln["ms_per_step"] / 5) < 1e-6
    assert ln["value"] == pytest.approx(5 * 1024 / (ln["ms_per_step"] * 1e-3))
