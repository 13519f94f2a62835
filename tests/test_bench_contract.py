"""The bench line the driver parses: the committed final line of the round (profiles/) and the reference-arm line carry every key of
the contract, with consistent values.  CPU only; bench.py itself needs a GPU and is exercised at round end."""

import json
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _latest(pattern):
    files = sorted((ROOT / "profiles").glob(pattern))
    assert files, pattern
    return json.loads(files[-1].read_text().strip().splitlines()[-1])


def test_own_arm_line_has_the_contract_keys():
    d = _latest("r02_bench_1gpu_32q*.json")
    base = json.loads((ROOT / "BASELINE.json").read_text())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["unit"] == "terms*amps/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "c128"
    assert d["warmup"] >= 3 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    assert "terms" in base["metric"] and "amps" in base["metric"]  # the metric BASELINE.json names
    # value = T * 2^n * steps / time
    cfg = d["config"]
    want = cfg["pauli_terms"] * 2.0 ** cfg["n_qubits"] / (d["ms_per_step"] * 1e-3)
    assert abs(d["value"] - want) <= 1e-6 * want
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 16 * 2 ** cfg["n_qubits"] and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.001
    assert e["energy_matches_device"] is True and e["call"].startswith("qfe_expectation") and e["host_memory"] == "pinned"
    assert d["gpu_launches"] == d["steps"] * (cfg["sweeps_per_step"] + 1)  # one launch per sweep + one reduction per step
    # the CPU baseline reports the matrix build and the evaluation separately, and both rates
    assert c["build_s"] > 0 and c["eval_s"] > 0 and c["value"] == c["value_cached_matrix"] > c["value_rebuild_per_step"] > 0
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_multi_gpu_lines_of_the_round():
    """The 2-, 4- and 8-GPU lines with the final kernel (builder runs, short step counts): whole-job value, weak scaling, NCCL exchanges
    counted, and the same energy as one GPU where the register is the same (SURVEY.md 8d iv)."""
    one = _latest("r02_bench_1gpu_32q*.json")
    for pattern, n_gpus, n_qubits in (("r02_bench_2gpu_32q_v2*.json", 2, 32), ("r02_bench_4gpu_33q_v2*.json", 4, 33), ("r02_bench_8gpu_34q_v2*.json", 8, 34)):
        d = _latest(pattern)
        cfg = d["config"]
        assert d["n_gpus"] == n_gpus and cfg["n_qubits"] == n_qubits and d["scaling"] == "weak" and d["dtype"] == "c128"
        want = cfg["pauli_terms"] * 2.0 ** n_qubits / (d["ms_per_step"] * 1e-3)
        assert abs(d["value"] - want) <= 1e-6 * want
        assert d["value"] > 0.9 * n_gpus * one["value"]  # at least 90 % of linear in the builder's own runs
        assert d["e2e"]["h2d_bytes_per_step"] == 16 * 2 ** n_qubits and d["exchanges_per_step"] == n_gpus - 1 and d["e2e"]["energy_matches_device"] is True
        if n_qubits == one["config"]["n_qubits"]:
            assert abs(d["energy"] - one["energy"]) <= 1e-12 * abs(one["energy"])
    # the 8-GPU run of the middle of the round (first-generation kernel) evaluated the same state: same energy
    mid, new = _latest("r02_bench_8gpu_34q_steps2.json"), _latest("r02_bench_8gpu_34q_v2*.json")
    assert abs(mid["energy"] - new["energy"]) <= 1e-12 * abs(new["energy"])


def test_reference_arm_line():
    d = _latest("r02_bench_reference_arm*.json")
    assert d["impl"] == "reference" and d["unit"] == "terms*amps/s" and d["value"] > 0
    assert "sample_n_qubits" in d["config"] and "BOUNDED SAMPLE" in d["config"]["workload"]  # the config names what actually ran
    assert d["cpu_baseline"]["build_s"] > 0 and d["cpu_baseline"]["value_rebuild_per_step"] < d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
