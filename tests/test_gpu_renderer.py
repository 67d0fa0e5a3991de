"""render_gaussians end to end on the GPU against the oracle pipeline (projection -> SH -> tile map ->
rasterize) and torch autograd through the torch restatements for the 3D parameter gradients."""
import pytest
import torch

import oracle
from oracle import torch_ref
from taichi_gaussian_rasterizer_b200 import RasterConfig, Rendering, render_gaussians
from util import GRAD_REL_L2, IMAGE_REL_L2, rel_l2, scene3d

pytestmark = pytest.mark.gpu


def oracle_render(g, cam, cfg, use_sh, render_depth, out):
  """CPU reference pipeline with autograd (taichi_splatting/renderer.py:133-231): torch restatements for the per
  point stages, the C++ oracle for tile mapping and rasterization.  The visible set comes from the C++ oracle
  (bit-exact contract with the CUDA projection); the tile map is built by the oracle from the SAME packed
  gaussians / sort depths the CUDA path produced, so both sides blend identical lists."""
  p_orc, d_orc, idx = oracle.projection_forward(*[t.detach() for t in g.shape_tensors()], cam.T_camera_world,
                                                cam.projection, cam.image_size, cam.depth_range, cfg.blur_cov,
                                                cfg.clamp_margin, cfg.alpha_threshold)
  assert torch.equal(out.points_in_view.cpu(), idx), "visible set differs from the oracle"
  assert torch.equal(out.gaussians2d.detach().cpu().view(torch.int32), p_orc.view(torch.int32))
  pts_all, depth_all = torch_ref.project_all(*g.shape_tensors(), cam.T_camera_world, cam.projection, cam.image_size,
                                             cfg.blur_cov, cfg.clamp_margin)
  pts, depth = pts_all[idx], depth_all[idx]
  if use_sh:
    feats = torch_ref.evaluate_sh_at(g.feature, g.position.detach(), idx, cam.camera_position)
  else:
    feats = g.feature[idx]
  if render_depth:
    feats = torch.cat([depth, depth ** 2, feats], dim=1)
  ndc = torch_ref.ndc_depth(out.point_depth.detach().cpu(), cam.near_plane, cam.far_plane)
  o2p, ranges = oracle.map_to_tiles(out.gaussians2d.detach().cpu(), ndc, cam.image_size, cfg)
  raster = oracle.rasterize_with_tiles(pts, feats, o2p, ranges.view(-1, 2), cam.image_size, cfg)
  return raster, idx, o2p


@pytest.mark.parametrize("use_sh,render_depth", [(False, False), (True, False), (True, True)])
def test_render_gaussians_vs_oracle(cuda_device, use_sh, render_depth):
  cfg = RasterConfig(compute_visibility=True, compute_point_heuristic=True)
  g, cam = scene3d(3, 6000, image_size=(320, 240), scale_factor=0.5, sh_degree=3 if use_sh else None)
  gd, cd = g.to(device=cuda_device), cam.to(device=cuda_device)
  gd.requires_grad_(True)
  gr = g.clone().requires_grad_(True)

  out = render_gaussians(gd, cd, cfg, use_sh=use_sh, render_depth=render_depth)
  ref, idx_ref, o2p_ref = oracle_render(gr, cam, cfg, use_sh, render_depth, out)
  assert isinstance(out, Rendering)
  full = out.image
  ref_image = ref.image[..., 2:] if render_depth else ref.image
  assert rel_l2(full, ref_image) < 5 * IMAGE_REL_L2
  assert rel_l2(out.image_weight, ref.image_weight) < 5 * IMAGE_REL_L2
  if render_depth:
    assert out.depth.shape == (240, 320) and out.depth_var.shape == (240, 320)
    d_ref = ref.image[..., 0] / (ref.image_weight + 1e-6)
    assert rel_l2(out.depth, d_ref) < 1e-4

  torch.manual_seed(1)
  gi = torch.rand_like(ref_image)
  (out.image * gi.to(cuda_device)).sum().backward()
  (ref_image * gi).sum().backward()
  for name in ("position", "log_scaling", "rotation", "alpha_logit", "feature"):
    a, b = getattr(gd, name).grad, getattr(gr, name).grad
    assert a is not None and rel_l2(a, b) < 5 * GRAD_REL_L2, f"{name}: {rel_l2(a, b)}"
  assert rel_l2(out.point_visibility, ref.visibility) < 5 * GRAD_REL_L2
  assert out.point_heuristic.shape == (idx_ref.shape[0], 2)
  assert out.split_score.shape == out.prune_cost.shape


def test_median_depth_and_depth16(cuda_device):
  cfg = RasterConfig()
  g, cam = scene3d(5, 4000, image_size=(256, 192), scale_factor=0.5)
  out = render_gaussians(g.to(device=cuda_device), cam.to(device=cuda_device), cfg, render_median_depth=True,
                         use_depth16=True)
  assert out.median_depth.shape == (192, 256)
  assert float(out.median_depth.min()) >= 0
  assert out.ndc_median_depth.shape == (192, 256)
