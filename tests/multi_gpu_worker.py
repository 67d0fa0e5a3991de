#!/usr/bin/env python
"""Parity of the view-parallel step on real GPUs (NCCL): every rank renders its share of a batch of views with
GradientBucket.fused_accumulation (deferred SH gradient, batched SH colours) and the ranks all-reduce once; rank 0 then
renders ALL views alone with plain autograd accumulation and compares.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_worker.py

Launched by tests/test_gpu_multi.py (pytest -m gpu; skipped on boxes with fewer than two GPUs).  Exit code 0 = parity.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from taichi_gaussian_rasterizer_b200 import RasterConfig, evaluate_sh_views, render_gaussians  # noqa: E402
from taichi_gaussian_rasterizer_b200.distributed import GradientBucket, partition_views  # noqa: E402
from taichi_gaussian_rasterizer_b200.synthetic import random_3d_gaussians, random_camera  # noqa: E402
from taichi_gaussian_rasterizer_b200.torch_lib.projection import join_rt, quat_to_mat  # noqa: E402


def main():
  rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
  torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
  dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
  dist.init_process_group("nccl", device_id=dev)
  torch.manual_seed(7)
  cam = random_camera(image_size=(320, 240))
  g_cpu = random_3d_gaussians(30_000, cam, scale_factor=0.8, sh_degree=3)
  views = 3 * world
  cams = [cam]
  for k in range(views - 1):
    q = torch.tensor([0.01 * (k % 4 + 1), -0.02 + 0.004 * k, 0.005, 1.0])
    cams.append(cam.transformed(join_rt(quat_to_mat(q / q.norm()), torch.tensor([0.02, 0.002 * k, -0.01]))))
  cams = [c.to(device=dev) for c in cams]
  cfg = RasterConfig()

  def grads(view_ids, fused, early=False, symmetric=False):
    g = g_cpu.to(device=dev)
    g.requires_grad_(True)
    bucket = GradientBucket([g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature], symmetric=symmetric)
    if symmetric and bucket.reducer is None:
      return None   # no multicast mapping on this box
    if fused:
      with bucket.fused_accumulation():
        bucket.zero_()
        mine = [cams[i] for i in view_ids]
        colors = evaluate_sh_views(g.feature, g.position, [c.camera_position for c in mine])
        for vi, (c, col) in enumerate(zip(mine, colors)):
          if early and vi == len(mine) - 1:
            bucket.reduce_early()   # SH slices reduced under the last view; its colour gradients are all-gathered
          render_gaussians(g, c, cfg, use_sh=True, sh_colors=col).image.square().mean().backward()
        bucket.all_reduce()
    else:
      for i in view_ids:
        render_gaussians(g, cams[i], cfg, use_sh=True).image.square().mean().backward()
    return bucket.flat.clone()

  mine = partition_views(views, rank, world)
  reduced = {"all-reduce at the end": grads(mine, fused=True), "reduce_early + gather": grads(mine, fused=True, early=True),
             # the bucket in symmetric memory, summed inside the NVSwitch by gs_multimem_all_reduce; with reduce_early the
             # last view's staged colour gradients are read from the peers' memory by the flush kernel itself
             "multimem all-reduce at the end": grads(mine, fused=True, symmetric=True),
             "multimem reduce_early + peer flush": grads(mine, fused=True, early=True, symmetric=True)}
  reduced = {k: v for k, v in reduced.items() if v is not None}
  if rank == 0:
    print(f"world {world}: variants {sorted(reduced)}")
  ok = True
  if rank == 0:
    ref = grads(list(range(views)), fused=False)
    for what, flat in reduced.items():
      err = ((flat.double() - ref.double()).norm() / ref.double().norm()).item()
      good = err < 1e-5 and ref.abs().sum().item() > 0
      ok = ok and good
      print(f"world {world}: {what}: gradient of {views} views vs single-GPU sum: rel l2 {err:.2e} -> {'OK' if good else 'FAIL'}")
  dist.barrier()
  dist.destroy_process_group()
  sys.exit(0 if ok else 1)


if __name__ == "__main__":
  main()
