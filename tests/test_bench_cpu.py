"""bench.py pieces that run without a GPU: the reference arm's JSON line (the contract keys the driver reads), the FLOP
accounting and the helpers around the timed region."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("collocation-points/sec") and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_flop_accounting_and_helpers():
    import bench
    import workloads
    from helpers import product_namespace
    from neurodiffeq_b200.tracing import TracedProblem
    from neurodiffeq_b200.engine import pad_scheme, combine_seconds
    wl = workloads.build(product_namespace(), "c2")
    assert wl.flops_fwdjet == 82816                                   # 256 + 10 * 8256
    tp = TracedProblem(wl.make_nets(), wl.make_conditions(), wl.diff_eqs, 2, pad_scheme=pad_scheme,
                       combine_seconds=combine_seconds)
    assert tp.n_channels == 4 and bench.executed_flops(wl, tp) == 256 + 8 * 8256   # combined channel: 4 of 5 channels run
    bf16, hbm, src = bench.load_peaks()
    assert bf16 > 100 and hbm > 1000 and isinstance(src, str)
    clk = bench.ClockSampler(0)
    with clk:
        pass
    s = clk.summary()
    assert "sm_mhz" in s and "reasons" in s
    from neurodiffeq_b200 import generators as G
    assert bench._survey_generator("c3", 65536, G).size == 65536
    with pytest.raises(KeyError):
        bench._survey_generator("zz", 4, G)
