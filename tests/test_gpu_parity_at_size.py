"""Parity with the oracle AT THE SIZES BASELINE.json NAMES: the workload the metric is quoted on (3 M gaussians,
SH degree 3, 2048x1365) and configurations 2-5 at full size, stage by stage through the public operators:

  project_to_image      visible set identical, packed gaussians and depths BIT-identical to the f32 oracle
  evaluate_sh_at        colours within 1e-5 relative L2
  map_to_tiles          overlap_to_point and tile_ranges BIT-identical
  rasterize_with_tiles  image / image_weight within 1e-5, visibility / gradients within 1e-4 (point heuristics 1e-3:
                        sums of absolute values and squares of the same per pixel terms), against the f32 oracle on the
                        very same packed gaussians, features, lists and image gradient
  render_gaussians      the composed entry point returns the same image as the staged operators
  projection backward   3D parameter gradients against the f32 restatement of the reverse sweep on the well
                        conditioned gaussians (see tests/test_gpu_projection.py for the conditioning argument)

The reference's own coverage of the full path at scale is tests/test_benchmarks.py:8-22 (1-2 M gaussians, runs only).
The oracle needs 10-40 s per configuration on 8 host cores.  Every measured error is appended to
gpurun_out/parity_at_size.jsonl (benchmarks/parity_report.py turns that into PARITY.md)."""
import json
import os
import time
from pathlib import Path

import pytest
import torch

import oracle
from oracle import torch_ref
from taichi_gaussian_rasterizer_b200 import (RasterConfig, evaluate_sh_at, map_to_tiles, rasterize_with_tiles,
                                             render_gaussians)
from taichi_gaussian_rasterizer_b200.perspective.projection import project_to_image
from taichi_gaussian_rasterizer_b200.synthetic import baseline_scene
from taichi_gaussian_rasterizer_b200.torch_lib.projection import ndc_depth
from util import GRAD_REL_L2, IMAGE_REL_L2, rel_l2

pytestmark = pytest.mark.gpu

REPORT = Path(__file__).resolve().parents[1] / "gpurun_out" / "parity_at_size.jsonl"
COND_MIN = 0.02   # conditioning mask of the projection backward comparison (fraction kept is reported)


def _record(row):
  try:
    REPORT.parent.mkdir(parents=True, exist_ok=True)
    with open(REPORT, "a") as f:
      f.write(json.dumps(row) + "\n")
  except OSError:
    pass


@pytest.mark.parametrize("name", ["bench", "c2", "c3", "c4", "c5"])
def test_full_size_parity_with_oracle(cuda_device, name):
  oracle.set_num_threads(len(os.sched_getaffinity(0)))
  g, cam, spec = baseline_scene(name)
  stats = bool(spec.get("stats"))
  use_sh = spec.get("sh_degree") is not None
  render_depth = bool(spec.get("render_depth"))
  cfg = RasterConfig(compute_visibility=stats, compute_point_heuristic=stats)
  size = cam.image_size
  row = {"config": name, "N": spec["n"], "image_size": list(size), "stats": stats}
  t0 = time.time()

  gd, cd = g.to(device=cuda_device), cam.to(device=cuda_device)
  gd.requires_grad_(True)

  # ---- projection: visible set and packed gaussians bit-identical
  g2d, depth, idx = project_to_image(gd, cd, cfg)
  p_orc, d_orc, idx_orc = oracle.project_to_image(g, cam, cfg)
  assert torch.equal(idx.cpu(), idx_orc), "visible set differs from the oracle"
  assert torch.equal(g2d.detach().cpu().view(torch.int32), p_orc.view(torch.int32)), "packed gaussians not bit-identical"
  assert torch.equal(depth.detach().cpu().view(torch.int32), d_orc.view(torch.int32)), "depths not bit-identical"
  row["V"] = int(idx.shape[0])

  # ---- colours
  if use_sh:
    feats = evaluate_sh_at(gd.feature, gd.position.detach(), idx, cd.camera_position)
    f_orc = oracle.evaluate_sh_at(g.feature, g.position, idx_orc, cam.camera_position)
    row["sh_colours"] = rel_l2(feats, f_orc)
    assert row["sh_colours"] < IMAGE_REL_L2
  else:
    feats = gd.feature[idx]
  if render_depth:
    feats = torch.cat([depth, depth ** 2, feats], dim=1)
  F = feats.shape[1]
  row["F"] = F

  # ---- tile map: bit-identical lists and ranges
  # The sort depth is the NDC depth torch computes ON THE DEVICE, as in the reference (torch_lib/projection.py:120-123
  # under torch.compile); torch's CPU kernels round the same expression differently in the last bit, so the oracle is
  # handed the device's bits (what the reference's tile mapper would see), not a host re-evaluation.
  ndc_gpu = ndc_depth(depth.detach(), cam.near_plane, cam.far_plane)
  ndc = ndc_gpu.cpu()
  assert rel_l2(ndc, torch_ref.ndc_depth(d_orc, cam.near_plane, cam.far_plane)) < 1e-6
  o2p, ranges = map_to_tiles(g2d.detach(), ndc_gpu, size, cfg)
  o2p_orc, ranges_orc = oracle.map_to_tiles(p_orc, ndc, size, cfg)
  assert torch.equal(o2p.cpu(), o2p_orc), "overlap_to_point differs from the oracle"
  assert torch.equal(ranges.cpu(), ranges_orc), "tile_ranges differ from the oracle"
  counts = (ranges_orc[..., 1] - ranges_orc[..., 0]).view(-1)
  row.update(K=int(o2p.shape[0]), K_per_tile_max=int(counts.max()), K_per_tile_mean=float(counts.float().mean()))

  # ---- rasterizer forward / backward on leaf copies of the very same inputs
  g2d_leaf = g2d.detach().clone().requires_grad_(True)
  f_leaf = feats.detach().clone().requires_grad_(True)
  out = rasterize_with_tiles(g2d_leaf, f_leaf, o2p, ranges.view(-1, 2), size, cfg)
  torch.manual_seed(11)
  gi = torch.rand(size[1], size[0], F) - 0.3
  (out.image * gi.to(cuda_device)).sum().backward()

  f_cpu = f_leaf.detach().cpu()
  img, w, vis = oracle.raster_forward(p_orc, f_cpu, o2p_orc, ranges_orc.view(-1, 2), size, cfg)
  heur = torch.zeros(p_orc.shape[0], 2)
  gp, gf = oracle.raster_backward(p_orc, f_cpu, o2p_orc, ranges_orc.view(-1, 2), size, cfg, img, gi, True, True, heur)
  row["image"] = rel_l2(out.image, img)
  row["image_weight"] = rel_l2(out.image_weight, w)
  row["grad_gaussians2d"] = rel_l2(g2d_leaf.grad, gp)
  row["grad_features"] = rel_l2(f_leaf.grad, gf)
  if stats:
    row["visibility"] = rel_l2(out.visibility, vis)
    row["point_heuristic"] = rel_l2(out.point_heuristic, heur)

  # ---- composed entry point: same image
  with torch.no_grad():
    full = render_gaussians(gd, cd, cfg, use_sh=use_sh, render_depth=render_depth)
  ref_rgb = out.image[..., 2:] if render_depth else out.image
  row["render_gaussians_vs_staged"] = rel_l2(full.image, ref_rgb)

  # ---- projection backward at size: CUDA f32 against the f32 restatement, well conditioned gaussians
  torch.manual_seed(12)
  go_p = torch.randn(idx.shape[0], 7)
  go_z = torch.randn(idx.shape[0], 1)
  for t in gd.shape_tensors():
    t.grad = None
  ((g2d * go_p.to(cuda_device)).sum() + (depth * go_z.to(cuda_device)).sum()).backward()
  ref_grads, cond = oracle.projection_backward(*g.shape_tensors(), cam.T_camera_world, cam.projection, size, idx_orc,
                                               go_p, go_z, blur_cov=cfg.blur_cov, clamp_margin=cfg.clamp_margin)
  good = torch.zeros(spec["n"], dtype=torch.bool)
  good[idx_orc[cond.min(dim=1).values > COND_MIN]] = True
  row["cond_fraction"] = float(good.sum()) / max(int(idx.shape[0]), 1)
  for k in ("position", "log_scaling", "rotation", "alpha_logit"):
    got = getattr(gd, k).grad.cpu()
    row[f"grad3d_{k}"] = rel_l2(got[good], ref_grads[k][good])
    row[f"grad3d_{k}_all"] = rel_l2(got, ref_grads[k])
  row["seconds"] = round(time.time() - t0, 1)
  _record(row)
  print(json.dumps(row))

  assert row["image"] < IMAGE_REL_L2 and row["image_weight"] < IMAGE_REL_L2, row
  assert row["grad_gaussians2d"] < GRAD_REL_L2 and row["grad_features"] < GRAD_REL_L2, row
  if stats:
    assert row["visibility"] < GRAD_REL_L2 and row["point_heuristic"] < 10 * GRAD_REL_L2, row
  assert row["render_gaussians_vs_staged"] < 1e-6, row
  assert row["cond_fraction"] > 0.85, row
  for k in ("position", "log_scaling", "rotation", "alpha_logit"):
    assert row[f"grad3d_{k}"] < GRAD_REL_L2, row
