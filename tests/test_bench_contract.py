"""CPU: bench.py's reference arm honours the driver's JSON contract (the CUDA arm needs a GPU and is exercised on the B200 box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(tmp_path):
    env = dict(os.environ, SPB_BENCH_WORKLOAD="dev", SPUMONI_B200_CACHE=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bases/sec PML-mode classification" and d["unit"] == "bases/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "spumoni_ref"))  # the reference's own code when it was built
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert ("oracle_port" in d) == have_ref
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_cuda_arm_refuses_to_run_without_a_gpu(tmp_path):
    """No CPU fallback: on a box without CUDA the product arm exits with an error instead of measuring something else."""
    import torch
    if torch.cuda.is_available():
        return
    env = dict(os.environ, SPB_BENCH_WORKLOAD="dev", SPUMONI_B200_CACHE=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""
