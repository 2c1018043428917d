"""CPU: the measurement contract of bench.py -- the algorithmic FLOP count behind `roofline.achieved` is the
formula of SURVEY.md 8(d), the reference arm prints a well-formed line, and the product arm refuses to run
without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_algorithmic_flops_match_survey_8d():
    import bench
    H, I = 768, 3072
    f_lin = 2 * (4 * H * H + 2 * H * I)
    assert f_lin == 14_155_776  # F_lin per token per layer
    enc = lambda L: 6 * (L * f_lin + 4 * L * L * H)
    per_pair = 2 * 50 * 2054 * H + enc(40) + enc(70) + 2 * enc(90) + 4 * H * H + 2 * 2 * H * H
    assert abs(per_pair / 1e9 - 25.21) < 0.01  # 25.21 GFLOP per pair forward (n_mul = 2)
    assert bench.flops_per_step(1, 40, 20, 50, 0, 0) == 3.0 * per_pair  # training step = 3 x forward
    # MLM head: 48.06 MFLOP per masked token
    head = (bench.flops_per_step(1, 40, 20, 50, 1, 0) - bench.flops_per_step(1, 40, 20, 50, 0, 0)) / 3.0
    assert abs(head / 1e6 - 48.06) < 0.01
    # the bench workload: ~19.7 TFLOP per step and GPU at batch 256 (SURVEY: 19.6)
    assert 19.5 < bench.flops_per_step(256, 40, 20, 50, 1344, 768) / 1e12 < 19.8


def test_product_arm_needs_a_gpu_and_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "needs a GPU" in (r.stderr + r.stdout)


def test_reference_arm_line_is_well_formed():
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["higher_is_better"] is True and line["gpu_launches"] == 0
