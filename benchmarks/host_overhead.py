#!/usr/bin/env python
"""Host-side cost of one render_gaussians forward + backward: a scene small enough that the GPU work is shorter than
the Python / launch path (20 k gaussians, 640x480), timed per frame and profiled with cProfile.

  python benchmarks/host_overhead.py [--frames 300] [--profile]
"""
import argparse
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from taichi_gaussian_rasterizer_b200 import RasterConfig, render_gaussians  # noqa: E402
from taichi_gaussian_rasterizer_b200.synthetic import random_3d_gaussians, random_camera  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--frames", type=int, default=300)
  ap.add_argument("--profile", action="store_true")
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  torch.manual_seed(0)
  cam = random_camera(image_size=(640, 480))
  g = random_3d_gaussians(20_000, cam, scale_factor=1.0, sh_degree=3).to(device=dev)
  g.requires_grad_(True)
  cam = cam.to(device=dev)
  cfg = RasterConfig()

  def frame():
    for t in (g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature):
      t.grad = None
    r = render_gaussians(g, cam, cfg, use_sh=True)
    r.image.mean().backward()

  for _ in range(20):
    frame()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(args.frames):
    frame()
  torch.cuda.synchronize()
  print(f"{(time.perf_counter() - t0) / args.frames * 1e3:.3f} ms per frame (host bound: 20 k gaussians, 640x480)")
  if args.profile:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(args.frames):
      frame()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
  main()
