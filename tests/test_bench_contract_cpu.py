"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port on the host cores) keeps the driver's contract -- exactly one
JSON line on stdout, printed by rank 0 only when launched under torch.distributed.run with 2 processes, ranks > 0 exit 0 without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_single_process_prints_one_json_line():
    lines = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-res", "32", "--gpus", "1"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "projection_images_steps_per_sec_1024" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["steps"] == 1 and d["warmup"] == 1


def test_reference_arm_under_torchrun_world2_prints_on_rank0_only():
    lines = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                  "--master-port", "29631", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-res", "32"])
    js = [l for l in lines if l.startswith("{")]
    assert len(js) == 1
    d = json.loads(js[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2
