"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the
keys the driver reads, our arm refuses to run without a CUDA device (no CPU fallback), and the parity metric of
euler2d_kokkos_b200/parity.py behaves as documented."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    if "unavailable" in d:
        pytest.skip(d["unavailable"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["unit"] == "Mcell-updates/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_our_arm_fails_loudly_without_a_gpu():
    out = run_bench("--steps", "1", "--no-cpu-baseline", timeout=300)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")], "no number without a GPU"
    assert "no CUDA device" in (out.stderr + out.stdout)


def test_parity_metric():
    from euler2d_kokkos_b200.parity import max_deviation, state_deviation

    rng = np.random.default_rng(0)
    b = np.empty((4, 12, 10))
    b[0] = rng.uniform(0.5, 2.0, b[0].shape)           # rho
    b[1] = rng.uniform(2.0, 5.0, b[1].shape)           # E
    b[2] = rng.uniform(-1.0, 1.0, b[2].shape)          # mx
    b[3] = 0.0                                         # my vanishes by symmetry
    assert max_deviation(b, b) == 0.0
    a = b.copy()
    a[0] *= 1 + 1e-13
    a[3] += 1e-16                                      # round-off in the vanishing component
    rows = {n: (l1, li) for n, l1, li in state_deviation(a, b)}
    assert 0.9e-13 < rows["rho"][0] < 1.1e-13 and 0.9e-13 < rows["rho"][1] < 1.1e-13
    assert rows["E"] == (0.0, 0.0) and rows["mx"] == (0.0, 0.0)
    # measured against sqrt(2 rho E) (>= 1.4 here), not against its own (zero) norm
    assert 0 < rows["my"][1] < 1e-16 and rows["my"][0] < 1e-16
