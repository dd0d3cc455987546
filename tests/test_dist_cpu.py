"""N > 1 host-side logic on CPU: two gloo processes (torchrun, 127.0.0.1)."""
import os
import socket
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_plumbing():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "dist_worker_cpu.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_reference_arm_only_rank0_prints(tmp_path):
    """bench.py --impl reference under a multi-rank launch: rank 0 alone runs and prints"""
    cmd = [sys.executable, os.path.join(os.path.dirname(HERE), "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--k", "8", "--layers", "2"]
    out1 = subprocess.run(cmd, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), capture_output=True, text=True,
                          timeout=300)
    assert out1.returncode == 0 and out1.stdout.strip() == ""
    out0 = subprocess.run(cmd, env=dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"), capture_output=True, text=True,
                          timeout=300)
    assert out0.returncode == 0, out0.stderr[-2000:]
    import json
    line = json.loads(out0.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gkr_prove_ms" and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
