"""bench.py's reference arm runs on the CPU: check the JSON contract of its one stdout line here (the GPU arm prints
the same keys plus roofline / clocks and is exercised on the B200 box)."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_all_host_threads():
    env = dict(os.environ, OMP_NUM_THREADS="1")            # what torchrun exports to its workers
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "matrices/s"
    assert d["metric"].startswith("weight-matrices factorised/sec")
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "modules", "svd_linear.py"))
    assert cb["kind"] == ("reference" if staged else "port") and cb["value"] == d["value"] and "4096x4096" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0))     # not the single thread OMP_NUM_THREADS=1 asks for
    assert "workload" in d["config"] and "model" not in d["config"]
    # the GPU arm prints the same `config` dict (bench.svd_config); run-dependent details live under `run`
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.svd_config(bench._default_batch_static(), 1843)


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
