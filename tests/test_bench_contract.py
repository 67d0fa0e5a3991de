"""bench.py prints ONE JSON line with the contract's keys: a small configuration on the GPU (`-m gpu`) and the CPU
reference arm on a tiny sample (CPU suite)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args, timeout=600):
  r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                     text=True, timeout=timeout, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
  assert len(lines) == 1, f"expected exactly one JSON line, got {len(lines)}"
  return json.loads(lines[0])


@pytest.mark.gpu
def test_bench_line_has_the_contract_keys(cuda_device):
  d = run_bench("--steps", "2", "--warmup", "3", "--views-per-rank", "2", "--num-gaussians", "200000",
                "--image-size", "640", "480", "--cpu-budget", "1")
  assert COMMON <= set(d) and {"roofline", "cpu_baseline", "clocks", "gpu_launches", "stage_ms_per_frame"} <= set(d)
  assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
  assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 4
  r = d["roofline"]
  assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
  assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
  assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
  assert "workload" in d["config"] and "model" not in d["config"]
  # the timed step is the CUDA-graph replay of the read-back free path, checked against the eager default path
  assert d["config"]["cuda_graph"] is True, d["config"].get("cuda_graph_error")
  assert d["config"]["graph_vs_eager_grad_rel_l2"] < 1e-5
  assert d["config"]["overlap_total_max"] <= d["config"]["overlap_capacity"]


def test_reference_arm_line(monkeypatch):
  d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--num-gaussians", "80000",
                "--image-size", "320", "240")
  assert d["impl"] == "reference" and COMMON <= set(d)
  assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
