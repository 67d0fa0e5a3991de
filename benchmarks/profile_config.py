#!/usr/bin/env python
"""torch.profiler kernel table of one BASELINE configuration step (where the time outside libgsplat's kernels goes).

  python benchmarks/profile_config.py c4 [--top 25]
"""
import argparse
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "benchmarks"))
sys.path.insert(0, str(ROOT))
import configs as cfgs  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("config")
  ap.add_argument("--top", type=int, default=25)
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  step = (cfgs.CONFIGS.get(args.config) or cfgs.EXTRA[args.config])(dev)
  for _ in range(3):
    step()
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
      step()
    torch.cuda.synchronize()
  rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
  total = sum(e.device_time_total for e in rows)
  print(f"{args.config}: {total / 3e3:.3f} ms of device time per step")
  for e in rows[:args.top]:
    print(f"{e.device_time_total / 3e3:8.3f} ms  x{e.count / 3:5.1f}  {e.key[:120]}")


if __name__ == "__main__":
  main()
