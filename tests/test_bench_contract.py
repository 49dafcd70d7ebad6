"""The bench.py contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the
agreed keys (and non-zero ranks of a torchrun launch stay silent)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, KLB_BENCH_CPU_SECONDS="0.5", **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--gpus", env_extra.get("WORLD_SIZE", "1")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_json_line():
    lines = _run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "leapfrog_steps_per_sec" and d["unit"] == "leapfrog-steps/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "65536 chains x 1024" in d["config"]["workload"]
    # both arms print the SAME config dict (the reference arm runs on this arm's config): bench.job_config
    import importlib.util
    spec = importlib.util.spec_from_file_location("klb_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert d["config"] == bench.job_config("reference")
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": job_config(') == 2          # the reference arm and the CUDA arm
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "chains" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4


def test_reference_arm_other_ranks_are_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
