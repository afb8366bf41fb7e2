"""CPU checks of the bench / host-side contract: the reference arm prints the JSON line the driver expects, the NUMA
binding helper never raises, ranks other than 0 of the reference arm exit silently."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, B200VQA_CPU_SAMPLE_PAIRS="1", **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    out = _run({})
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "videos_per_sec_1080p_e2e" and d["unit"] == "videos/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"] == "1080p-10s"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == dict(value=d["value"], unit="videos/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_other_ranks_are_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bind_host_to_gpu_never_raises():
    from relax_vqa_b200.engine import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    cpus = bind_host_to_gpu(0)
    assert cpus is None or (isinstance(cpus, set) and cpus <= before)
    os.sched_setaffinity(0, before)
