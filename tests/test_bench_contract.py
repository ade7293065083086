"""bench.py prints ONE JSON line with the keys the driver parses, for every arm and workload."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _bench(*args, timeout=900):
    p = subprocess.run([sys.executable, "bench.py", *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, PYTHONPATH=ROOT))
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line: %r" % (lines,)
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    """The reference arm needs no GPU: the unmodified reference (or the oracle port) streamed on the host cores."""
    d = _bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--shape", "drow")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_cutout_sweep_reference_arm():
    d = _bench("--workload", "cutout", "--impl", "reference", "--steps", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["cpu_baseline"]["cores"] == 1 and d["value"] > 0


@pytest.mark.gpu
def test_stream_arm_prints_the_contract_line():
    d = _bench("--steps", "2", "--warmup", "3", "--sequences", "8", "--no-cpu-baseline")
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "parity_spot"} <= set(d)
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # the 1e-5 bar is on identical cutouts; `max_rel` also carries NumPy's float32 arctan (1-2 ulp, host-CPU specific), which can
    # flip a nearest-beam tap of an area row (tests/test_gpu_timed_config.py arbitrates that with float64)
    assert d["parity_spot"]["max_rel_network_on_same_cutouts"] <= 1e-5 and d["parity_spot"]["max_rel"] <= 2e-4
    assert d["parity_spot"]["mask_equal"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


@pytest.mark.gpu
def test_cutout_sweep_arm_prints_the_contract_line():
    d = _bench("--workload", "cutout", "--steps", "3", "--no-cpu-baseline")
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "sweep", "parity_spot"} <= set(d)
    assert [r["batch"] for r in d["sweep"]] == [1, 4, 16, 64, 256, 1024, 4096]
    assert d["parity_spot"]["max_rel_with_reference_half_angles"] <= 1e-5 and d["parity_spot"]["bit_equal_fraction"] >= 0.9999
    assert d["e2e"]["d2h_bytes_per_step"] == 4096 * 1091 * 56 * 4
