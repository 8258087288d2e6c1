"""The bench.py JSON contract, as far as it can be exercised without a GPU: the reference arm (`--impl reference`) runs
the CPU restatement and must print ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                       cwd=ROOT, env=e)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-refine", "4")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fish3d_cg_gmg_solve_mdof_per_s" and d["unit"] == "MDOF/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "MDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["value"] > 0


def test_reference_arm_under_torchrun_only_rank0_prints():
    assert run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-refine", "3", "--gpus", "2",
                     env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
