"""bench.py's side of the measurement contract that can be checked without a GPU: the reference arm's JSON line, the
helpers that label and size the B200 arm's line, and stdout hygiene."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("workload", ["fft", "windowed", "helfand"])
def test_reference_arm_line_has_every_contract_key(workload):
    sizes = {"fft": ("48", "800"), "windowed": ("4", "300"), "helfand": ("4", "300")}[workload]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                        "--steps", "1", "--warmup", "1", "--atoms", sizes[0], "--frames", sizes[1]],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                      # stdout is the JSON line and nothing else
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "atom-frames/s" and d["unit"] == "atom-frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]


def test_algorithmic_work_per_atom_frame_matches_the_survey():
    # 32 B (VACF) / 56 B (Helfand) per atom-frame; 491.5 flop at T = 10,000, 458.8 at T = 5,000
    assert bench.BYTES_PER_AF == {"fft": 32.0, "windowed": 32.0, "helfand": 56.0, "helfand_direct": 56.0, "helfand_fft": 56.0}
    # FP32 mode: float series in HBM (4 D bytes in), float64 per-particle rows (8 bytes out)
    assert bench.BYTES_PER_AF_FP32["fft"] == 20.0 and bench.BYTES_PER_AF_FP32["helfand"] == 32.0
    assert bench.flops_per_af("fft", 10000) == pytest.approx(491.52, abs=0.01)
    assert bench.flops_per_af("fft", 5000) == pytest.approx(458.75, abs=0.01)
    assert bench.flops_per_af("windowed", 2000) == 3 * 2001
    assert bench.flops_per_af("helfand_direct", 5000) == 1.5 * 3 * 4999
    assert bench.DEFAULTS["fft"] == (100_000, 10_000) and bench.DEFAULTS["windowed"] == (1_000, 2_000)
    assert bench.DEFAULTS["helfand"] == (10_000, 5_000)


def test_kernel_label_follows_the_plan_the_library_reports():
    assert "three-pass" in bench.k1_kernel_name({"radices": [20, 16, 16], "threads": 320})
    assert "float" in bench.k1_kernel_name({"radices": [20, 16, 16], "threads": 320}, "fp32")
    assert "general" in bench.k1_kernel_name({"radices": [8, 8, 4], "threads": 32})


def test_tiled_synthetic_trajectory_gives_every_particle_its_own_series():
    v = bench.synthetic_trajectory(50, 3 * bench.TILE_ATOMS + 5, seed=3, threads=2)
    assert v.shape == (50, 3 * bench.TILE_ATOMS + 5, 3) and v.dtype == np.float32
    a, b = v[:, 7], v[:, bench.TILE_ATOMS + 7]
    assert np.allclose(b, a * np.float32(1.0 + 1 / 64.0)) and not np.array_equal(a, b)
    assert np.array_equal(v[:, : bench.TILE_ATOMS], bench.synthetic_trajectory(50, bench.TILE_ATOMS, seed=3, threads=1))
    assert abs(bench.sum_of_squares(v, threads=3) - float((v.astype(np.float64) ** 2).sum())) < 1e-6 * float((v.astype(np.float64) ** 2).sum())


def test_numa_binding_degrades_to_a_note_without_nvml():
    before = os.sched_getaffinity(0)
    info = bench.bind_near_gpu(0)
    assert "node" in info and (info["node"] is None or isinstance(info["node"], int))
    if info["node"] is None:
        assert os.sched_getaffinity(0) == before and info["why"]


def test_seeded_fill_is_reproducible_and_chunk_independent():
    a, b = np.empty((3, 1000, 3), np.float32), np.empty((3, 1000, 3), np.float32)
    bench.fill_random_f32(a, seed=5, threads=1)
    bench.fill_random_f32(b, seed=5, threads=4)
    assert np.array_equal(a, b) and abs(float(a.mean())) < 0.1 and 0.9 < float(a.std()) < 1.1
