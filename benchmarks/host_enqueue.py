#!/usr/bin/env python
"""How long the HOST needs to enqueue one bench frame (Python + ctypes + torch launches), against the device time of the
frame: the ratio is the slack the host has before a step becomes launch bound (it shrinks when eight ranks share the
host cores).  python benchmarks/host_enqueue.py [--views 8]"""
import argparse
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from taichi_gaussian_rasterizer_b200 import RasterConfig, evaluate_sh_views, render_gaussians  # noqa: E402
from taichi_gaussian_rasterizer_b200.distributed import GradientBucket  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--views", type=int, default=8)
  ap.add_argument("--profile", action="store_true")
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  g_cpu, cams, _ = bench.build_scene("bench", 0, args.views)
  g = g_cpu.to(device=dev)
  g.requires_grad_(True)
  cams = [c.to(device=dev) for c in cams]
  w, h = cams[0].image_size
  targets = [torch.rand(h, w, 3, device=dev) for _ in cams]
  cfg = RasterConfig()
  bucket = GradientBucket([g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature])

  def step():
    with bucket.fused_accumulation():
      bucket.zero_()
      colors = evaluate_sh_views(g.feature, g.position, [c.camera_position for c in cams])
      for cam, tgt, col in zip(cams, targets, colors):
        r = render_gaussians(g, cam, cfg, use_sh=True, sh_colors=col)
        torch.nn.functional.l1_loss(r.image, tgt).backward()

  for _ in range(3):
    step()
  torch.cuda.synchronize()
  n = 5
  t0 = time.perf_counter()
  for _ in range(n):
    step()
  t1 = time.perf_counter()
  torch.cuda.synchronize()
  t2 = time.perf_counter()
  print(f"per frame: host returns after {(t1 - t0) / n / args.views * 1e3:.3f} ms, device done after "
        f"{(t2 - t0) / n / args.views * 1e3:.3f} ms (the two host read-backs per frame make the host wait for the device)")
  if args.profile:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
      step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
  main()
