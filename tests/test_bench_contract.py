"""The CPU-runnable half of bench.py's contract: `--impl reference` (the oracle port on the host cores, the one place
besides cpu_baseline where bench.py executes oracle/) prints one JSON line with the keys the driver reads, and ranks
other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                           "--warmup", "1", "--ref-envs", "1"], env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-800:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "AisleTurnEnv" in d["config"]["workload"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    from oracle.ref_loader import reference_available
    assert cb["kind"] == ("reference" if reference_available() else "port")      # the unmodified reference when importable
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"] and cb["port_value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_0_only():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_b200_arm_prints_the_contract_line():
    """A small run of the CUDA arm: every key of the measurement contract is there and self-consistent."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--envs", "2048", "--pool", "32", "--steps", "5",
                          "--warmup", "3", "--e2e-steps", "2", "--e2e-image-steps", "1", "--cpu-seconds", "0.5",
                          "--gen-envs", "0"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-800:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert "impl" not in d and d["metric"] == "env_steps_per_sec" and d["steps"] == 5 and d["warmup"] == 3
    assert abs(d["value"] - 2048 * 5 / (d["ms_per_step"] * 5e-3)) < 1e-6 * d["value"]
    assert d["gpu_launches"] == 5 * 3                      # move_kernel, reward_kernel, sparse egocentric kernel per step
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] is None      # ncu traffic is for 65 536 envs
    assert d["roofline_collision"]["kernel"].startswith("collision_thread_kernel") and d["roofline_collision"]["frac"] > 0
    for name in ("move_kernel", "reward_kernel", "ego_sparse_kernel"):
        assert d["kernels_ms"][name] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 2048 * 2 * 4 and e["d2h_bytes_per_step"] == 2048 * (8 + 1 + 48)
    im = d["e2e_images_to_host"]
    assert im["dense"]["d2h_bytes_per_step"] > 2048 * 117 * 133 > 4 * im["compact"]["d2h_bytes_per_step"] and im["compact"]["value"] > 0
    assert d["value_graph"]["value"] > 0 and d["value_200_steps"]["steps"] == 200 and d["strong"] is None     # (2048 envs: not the strong workload)
    cb = d["cpu_baseline"]
    from oracle.ref_loader import reference_available
    assert cb["kind"] == ("reference" if reference_available() else "port") and cb["cores"] == 1 and cb["value"] > 0 and cb["sample"]
    assert d["clocks"]["samples"] >= 1 and d["clocks"]["sm_mhz"] is not None
