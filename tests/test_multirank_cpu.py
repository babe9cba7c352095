"""N > 1 on the CPU: world_size-2 gloo runs of the host-side multi-rank logic (no GPU needed).

The data path has no collective except one all-reduce of T + 1 doubles per compute call; these tests
cover the launch contract of bench.py under torchrun and the shard -> sum -> all-reduce -> mean protocol itself."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

import oracle
from _util import assert_close_normwise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, script_args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port())] + script_args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_sharded_sum_allreduce_mean_equals_single_rank(tmp_path):
    T, N, seed = 400, 11, 5                                        # 11 particles over 2 ranks: ragged split 5 + 6
    r = _torchrun(2, [os.path.join(ROOT, "tests", "_gloo_worker.py"), str(tmp_path), str(T), str(N), str(seed)])
    assert r.returncode == 0, r.stderr[-2000:]
    vel = np.random.default_rng(seed).standard_normal((T, N, 3)).astype(np.float32).astype(np.float64)
    _, ref_ts = oracle.vacf_fft(vel)
    got = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    assert (int(got[0]["a0"]), int(got[0]["a1"]), int(got[1]["a0"]), int(got[1]["a1"])) == (0, 5, 5, 11)
    for g in got:
        assert g["count"] == N and g["tmax"] == 2.0
        assert_close_normwise(g["ts"], ref_ts, 1e-13, "sharded mean vs single-rank oracle")
    assert np.array_equal(got[0]["ts"], got[1]["ts"])              # every rank holds the same reduced timeseries


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                      "--atoms", "64", "--frames", "600"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "atom-frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_b200_arm_launches_its_ranks_and_fails_loudly_without_gpus():
    """`bench.py --gpus 2` outside torchrun starts its own two ranks; on a machine without GPUs every rank must stop with
    the backend's error (no CPU fallback) and no JSON line may appear."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["MASTER_PORT"] = str(_free_port())
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
                        "--atoms", "8", "--frames", "64"], cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    try:
        import ctypes
        have_gpu = ctypes.CDLL("libcuda.so.1").cuInit(0) == 0
    except OSError:
        have_gpu = False
    if have_gpu:
        return
    assert r.returncode != 0
    assert "BackendError" in r.stderr, r.stderr[-1500:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
