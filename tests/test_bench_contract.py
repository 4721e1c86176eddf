"""bench.py's reference arm runs on the CPU: check that it honours the driver's JSON contract
(one line, the metric/unit/config of BASELINE.json, cpu_baseline + e2e objects) on a tiny sample."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-n", "256"], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_json_contract():
    d = _run()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference"
    assert d["metric"].split(";")[0] == base["metric"].split(";")[0]
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs and prints the reference arm."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_b200_bench_line_has_every_contract_key():
    """The last bench line measured on the B200 (profiles/): the keys the driver and the judge read."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_g1_v2.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] > 0
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the in-situ figure: the dominant kernel cannot have taken longer than the step that contains it
    assert abs(r["achieved"] - r["algorithmic_flops_per_step"] / (d["ms_per_step"] * 1e-3) / 1e12) < 1e-6 * r["achieved"]
    assert r["serialised_step"]["gemm_share_of_step"] <= 1.0


def test_committed_8gpu_bench_line_carries_the_block_cyclic_record():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_g8_v2.json")))
    assert d["n_gpus"] == 8 and d["scaling"] == "weak"
    b = d["block_cyclic"]
    assert b["N"] == 131072 and b["n_gpus"] == 8 and b["grid"] == [4, 2]
    for key in ("factor_s", "sweep_s", "eval_s", "eval_tflops_per_gpu", "strong_scaling_efficiency", "agreement", "lml"):
        assert key in b, key
    assert b["agreement"]["lml_rel_diff_vs_single_gpu"] <= 1e-11 and b["agreement"]["grad_rel_diff_vs_single_gpu"] <= 1e-9
    assert 0.0 < b["strong_scaling_efficiency"] <= 1.0
