"""The bench line's contract (keys the driver and the judge read), checked on the CPU: the reference arm is run live with one
step, rank > 0 of a torchrun launch of that arm exits 0 without printing, and the committed B200 lines under ``profiles/`` are
validated field by field (they are what ``bench.py`` printed on the box; this guards the format, not the numbers)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config"}


def _last_json_line(text: str) -> dict:
    lines = [l for l in text.strip().splitlines() if l.startswith("{")]
    assert lines, text[-400:]
    return json.loads(lines[-1])


def _check_e2e(e2e: dict) -> None:
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e2e)
    assert e2e["value"] > 0


def test_reference_arm_prints_one_contract_line_and_other_ranks_stay_silent():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    assert len([l for l in out.stdout.splitlines() if l.startswith("{")]) == 1
    line = _last_json_line(out.stdout)
    assert BASE_KEYS <= set(line) and line["impl"] == "reference" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 1 and line["value"] > 0 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "100 000" in cb["sample"]
    _check_e2e(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # a non-zero rank of the torchrun launch of this arm: exit 0, nothing printed, no rendezvous attempted
    env2 = dict(env, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
    out2 = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                          cwd=ROOT, env=env2, capture_output=True, text=True, timeout=300)
    assert out2.returncode == 0 and out2.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


@pytest.mark.parametrize("name", ["r2_bench_final.json", "r2_bench_n2.json", "r2_bench_n4.json", "r2_bench_n8.json"])
def test_committed_rainbow_lines_carry_every_contract_field(name):
    line = _last_json_line(open(os.path.join(ROOT, "profiles", name)).read())
    assert BASE_KEYS <= set(line) and "impl" not in line
    assert line["metric"].startswith("population gradient-steps/sec") and line["unit"] == "steps/s"
    assert line["scaling"] == "strong" and line["dtype"] == "f32" and line["data"] == "synthetic" and line["warmup"] >= 3
    assert abs(line["value"] - 8 * line["steps"] / (line["ms_per_step"] * line["steps"] / 1e3)) / line["value"] < 1e-6
    assert line["gpu_launches"] > 0
    c = line["clocks"]
    assert c["samples"] > 0 and c["sm_mhz"] > 0.9 * c["sm_max_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert "l2" in line["config"] and "workload" in line["config"]
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert abs(r["achieved"] - r["alg_bytes_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) / r["achieved"] < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    assert 0 < r["tensor"]["frac"] < 1
    _check_e2e(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] < line["value"]                       # an e2e that repeats the device number is no e2e
    if line["n_gpus"] == 1:
        assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    else:
        t = line["tournament"]
        assert t["plans_identical"] is True and t["bytes_allgather"] > 0 and t["moved_agents"] >= 0


@pytest.mark.parametrize("name", ["r2_bench_td3.json", "r2_bench_ppo.json", "r2_bench_maddpg.json"])
def test_committed_next_row_lines(name):
    line = _last_json_line(open(os.path.join(ROOT, "profiles", name)).read())
    assert BASE_KEYS <= set(line) and line["value"] > 0 and "workload" in line["config"]
    _check_e2e(line["e2e"])
    assert line["cpu_baseline"]["value"] > 0 and line["roofline"]["frac"] > 0 and line["gpu_launches"] > 0
