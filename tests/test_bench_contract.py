"""CPU checks of the bench / host-side contract: the reference arm prints the JSON line the driver expects, the NUMA
binding helper never raises, ranks other than 0 of the reference arm exit silently."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, extra_args=()):
    env = dict(os.environ, B200VQA_CPU_SAMPLE_PAIRS="1", **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                           *extra_args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    out = _run({})
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "videos_per_sec_1080p" and d["unit"] == "videos/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"] == "1080p-10s"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == dict(value=d["value"], unit="videos/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_other_workloads():
    """The other BASELINE.json configs: same line, the workload's own metric name and config; the LSVQ mix draws its
    resolutions from the committed shape histogram."""
    sys.path.insert(0, ROOT)
    import bench
    out = _run({}, ("--workload", "540p-8s"))
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["metric"] == "videos_per_sec_540p" and d["config"] == bench.workload_config("540p-8s", 16, 1)
    specs = bench.global_specs("lsvq-mix", 96)
    assert specs == bench.global_specs("lsvq-mix", 96) and len({(h, w) for h, w, _ in specs}) > 5        # seeded, mixed
    assert all(h >= 64 and w >= 64 and 1 <= p <= 64 for h, w, p in specs)
    assert bench.global_specs("2160p-20s", 2) == [(2160, 3840, 43)] * 2
    assert bench.metric_name("lsvq-mix") == "videos_per_sec_lsvq_mix" and bench.metric_name("2160p-20s") == "videos_per_sec_2160p"


def test_reference_arm_other_ranks_are_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bind_host_to_gpu_never_raises():
    from relax_vqa_b200.engine import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    cpus = bind_host_to_gpu(0)
    assert cpus is None or (isinstance(cpus, set) and cpus <= before)
    os.sched_setaffinity(0, before)
