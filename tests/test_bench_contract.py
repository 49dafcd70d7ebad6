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


def test_cuda_arm_line_assembles_without_a_gpu():
    """The dict the CUDA arm prints is evaluated here with stand-in values for the measured quantities: a typo in that
    expression would otherwise only show on the GPU box, after the whole run."""
    import ast
    import importlib.util
    import types
    path = os.path.join(ROOT, "bench.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_gpu"][0]
    dicts = [n.value for n in ast.walk(fn) if isinstance(n, ast.Assign) and isinstance(n.value, ast.Dict)
             and any(isinstance(k, ast.Constant) and k.value == "metric" for k in n.value.keys)]
    assert len(dicts) == 1
    expr = ast.Expression(dicts[0])
    ast.fix_missing_locations(expr)
    spec = importlib.util.spec_from_file_location("klb_bench2", path)
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Plan:
        nv, regs_per_thread, blocks_per_sm = 16, 128, 2
    ns = dict(vars(bench))
    ns.update(value=2.19e9, world=1, args=types.SimpleNamespace(steps=10, warmup=3, e2e_serial=False), dev_ms=598.0,
              arith="reference", nloc=65536, gather=None, gather_note=None, plan=Plan(), acc_rate=0.99, wall_ms=600.0,
              fp64_achieved=1.1e13, fp64_peak=1.86e13, traffic={"dram_bytes_per_launch": 5.4e10, "nchains": 65536},
              traffic_src="x", kernel_ms=59.8, ops_launch=1, hbm_achieved=4400.0, hbm_peak=6552.0, bytes_launch=1, peak_src="m",
              cb={"value": 1}, lf_per_step=131072000, e2e_ms=61.0, h2d=1, d2h=1,
              e2e_full={"ms_per_step": 1000.0, "d2h_bytes_per_step": 1, "h2d_bytes_per_step": 1}, ess_sum=1.0, ess_min=1.0,
              ess_ms=29.0, parity={"bit_exact": True}, configs={}, fp64_stream=1.7e13, dmma_peak=3.7e13, launches_timed=20,
              clocks={"sm_mhz": 1965.0})
    out = eval(compile(expr, "bench_out", "eval"), ns)
    json.dumps(out)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity", "configs"):
        assert k in out, k
    assert out["config"] == bench.job_config("reference") and out["roofline"]["bound"] == "fp64_issue"
    for k in ("achieved", "peak", "unit", "frac", "traffic"):
        assert k in out["roofline"], k
    assert set(out["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
