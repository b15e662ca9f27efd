"""CPU: flat parameter layout invariants and the committed bench line's contract keys."""
import json
import os

import pytest


def test_flat_layout_northstar_counts_and_alignment():
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    spec = rainbow_spec((4, 84, 84), 6, channel_size=(32, 32), kernel_size=(8, 4), stride_size=(4, 2),
                        obs_low=0.0, obs_high=255.0, obs_u8=True)
    layout = FlatLayout(spec)
    assert layout.n_param_elems == 162730 and layout.n_eps_elems == 27429          # reference counts
    assert layout.n_params >= layout.n_param_elems and layout.n_eps >= layout.n_eps_elems
    end = {"param": 0, "eps": 0}
    for key, e in layout.entries.items():
        assert e.offset % 4 == 0, f"{key} is not 16-byte aligned"                   # float4 loads in the kernels
        n = 1
        for d in e.shape:
            n *= d
        assert e.offset >= end[e.buf], f"{key} overlaps its predecessor"
        end[e.buf] = e.offset + n
    assert end["param"] <= layout.n_params and end["eps"] <= layout.n_eps
    # reference state_dict names survive (checkpoints / mutations address parameters by them)
    assert "encoder.model.encoder_conv_layer_1.weight" in layout.entries
    assert "head_net.model.value_linear_layer_1.weight_mu" in layout.entries
    assert "head_net.advantage_net.advantage_linear_layer_output.weight_sigma" in layout.entries


@pytest.mark.parametrize("name", ["r1_bench_final.json", "r1_bench_final_n2.json", "r1_bench_reference_arm.json"])
def test_committed_bench_lines_keep_the_contract(name):
    path = os.path.join(os.path.dirname(__file__), "..", "profiles", name)
    if not os.path.exists(path):
        pytest.skip("profile not committed")
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e"):
        assert k in d, f"{name}: missing {k}"
    assert d["unit"] == "steps/s" and d["higher_is_better"] is True and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    if d.get("impl") == "reference":
        assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    else:
        assert d["gpu_launches"] > 0 and "clocks" in d
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        if d["n_gpus"] == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
