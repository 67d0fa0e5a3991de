#!/usr/bin/env python
"""What --use_fast_math on csrc/point_kernels.cu costs in accuracy and buys in time.

Runs the projection backward and the spherical-harmonics kernels of the bench workload (3 M gaussians, 2048x1365) with
the default library and with libgsplat_b200_precise.so (same sources, point_kernels.cu compiled without the flag,
`python -m taichi_gaussian_rasterizer_b200.csrc.build --precise-point-kernels`), each in its own process, and prints
one JSON line per build: relative L2 of the CUDA f32 gradients against the f32 restatement of the reverse sweep
(oracle.projection_backward<float>) and against its f64 instantiation, over all gaussians and over the well conditioned
ones, and the device time of the entry points.

  python benchmarks/fast_math_cost.py [--n 3000000]
"""
import argparse
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def child(variant, n):
  import torch
  from taichi_gaussian_rasterizer_b200 import _native
  if variant == "precise":
    _native.use_library(ROOT / "taichi_gaussian_rasterizer_b200" / "libgsplat_b200_precise.so")
  import oracle
  from taichi_gaussian_rasterizer_b200 import RasterConfig, evaluate_sh_at
  from taichi_gaussian_rasterizer_b200.perspective.projection import project_to_image
  from taichi_gaussian_rasterizer_b200.synthetic import baseline_scene

  dev = torch.device("cuda:0")
  g, cam, spec = baseline_scene("bench", n=n)
  cfg = RasterConfig()
  gd, cd = g.to(device=dev), cam.to(device=dev)
  gd.requires_grad_(True)
  torch.manual_seed(5)
  g2d, depth, idx = project_to_image(gd, cd, cfg)
  V = idx.shape[0]
  go_p, go_z, go_c = torch.randn(V, 7), torch.randn(V, 1), torch.randn(V, 3)
  go_pd, go_zd, go_cd = go_p.to(dev), go_z.to(dev), go_c.to(dev)

  def step():
    for t in gd.shape_tensors() + (gd.feature,):
      t.grad = None
    g2d, depth, idx = project_to_image(gd, cd, cfg)
    col = evaluate_sh_at(gd.feature, gd.position.detach(), idx, cd.camera_position)
    ((g2d * go_pd).sum() + (depth * go_zd).sum() + (col * go_cd).sum()).backward()

  for _ in range(3):
    step()
  torch.cuda.synchronize()
  timer = _native.set_stage_timer(_native.StageTimer())
  for _ in range(10):
    step()
  stage = {k: round(v[1] / v[0], 4) for k, v in timer.summary().items()}
  _native.set_stage_timer(None)

  args = (*g.shape_tensors(), cam.T_camera_world, cam.projection)
  r32, cond = oracle.projection_backward(*args, cam.image_size, idx.cpu(), go_p, go_z, blur_cov=cfg.blur_cov)
  r64, _ = oracle.projection_backward(*[a.double() for a in args], cam.image_size, idx.cpu(), go_p.double(),
                                      go_z.double(), blur_cov=cfg.blur_cov)
  good = torch.zeros(g.position.shape[0], dtype=torch.bool)
  good[idx.cpu()[cond.min(dim=1).values > 0.02]] = True

  def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()

  out = {"build": variant, "N": n, "V": V, "conditioned_fraction": float(good.sum()) / V, "ms": stage}
  for k in ("position", "log_scaling", "rotation", "alpha_logit"):
    got = getattr(gd, k).grad.cpu()
    out[k] = {"conditioned_vs_f32": rel(got[good], r32[k][good]), "conditioned_vs_f64": rel(got[good], r64[k][good]),
              "all_vs_f32": rel(got, r32[k]), "all_vs_f64": rel(got, r64[k])}
  print(json.dumps(out), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--n", type=int, default=3_000_000)
  ap.add_argument("--child", default=None)
  args = ap.parse_args()
  if args.child:
    child(args.child, args.n)
    return
  from taichi_gaussian_rasterizer_b200.csrc import build
  build.build(precise_point_kernels=True)
  for variant in ("fast", "precise"):
    subprocess.run([sys.executable, __file__, "--child", variant, "--n", str(args.n)], check=False)


if __name__ == "__main__":
  main()
