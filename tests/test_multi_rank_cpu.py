"""N > 1 host logic on CPU: world_size-2 gloo run (sharding + counter-RNG slicing) and the reference
arm of bench.py under torchrun (rank 0 alone prints, the other rank exits 0 without work)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script_args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_two_rank_sharding_matches_unsharded():
    p = _torchrun([os.path.join(ROOT, "tests", "dist_worker.py")], 29531)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "SHARD_OK" in p.stdout


def test_reference_arm_under_torchrun_prints_once():
    p = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"], 29532)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "GMAC/s" and j["value"] > 0 and j["n_gpus"] == 2
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["e2e"]["h2d_bytes_per_step"] == 0
