#!/usr/bin/env python
"""torch.profiler view of one BASELINE configuration step (benchmarks/configs.py): device time per kernel / ATen op and
host enqueue time per step — shows how much of a configuration is the rasterizer and how much is glue.

  python benchmarks/host_profile.py c4 [--rows 30]
"""
import argparse
import sys
import time
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "benchmarks"))
import configs  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("config")
  ap.add_argument("--rows", type=int, default=30)
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  step = (configs.CONFIGS.get(args.config) or configs.EXTRA[args.config])(dev)
  for _ in range(3):
    step()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(10):
    step()
  t1 = time.perf_counter()
  torch.cuda.synchronize()
  t2 = time.perf_counter()
  print(f"host enqueue per step {(t1 - t0) / 10 * 1e3:.3f} ms, with drain {(t2 - t0) / 10 * 1e3:.3f} ms")
  with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
      step()
    torch.cuda.synchronize()
  print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=args.rows, max_name_column_width=70))


if __name__ == "__main__":
  main()
