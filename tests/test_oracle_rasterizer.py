"""Oracle checks for the rasterizer (CPU): the relations the reference's tests pin, run on the oracle —
float64 gradcheck of the single tile setup (tests/test_rasterizer.py:30-90), visibility == feature
gradient under a unit image gradient (tests/test_visibility.py:34-63) — plus the stale-tail quirk
 and its index arithmetic."""
import numpy as np
import pytest
import torch

import oracle
from taichi_gaussian_rasterizer_b200 import RasterConfig
from util import rel_l2, scene2d


def single_tile_inputs(seed, config):
  torch.manual_seed(seed)
  n = int(torch.randint(1, 50, (1,)))
  channels = int(torch.randint(1, 4, (1,)))
  g, _, feat = scene2d(seed, n, (8, 8), channels=channels, alpha_range=(0.2, 0.8), dtype=torch.float64)
  o2p = torch.arange(0, n, dtype=torch.int32)
  ranges = torch.tensor([[0, n]], dtype=torch.int32)

  def render(mean, axis, sigma, alpha, colors):
    packed = torch.cat([mean, axis, sigma, alpha], dim=-1)
    return oracle.rasterize_with_tiles(packed, colors, o2p, ranges, (8, 8), config).image

  inputs = (g[:, 0:2], g[:, 2:4], g[:, 4:6], g[:, 6:7], feat)
  return tuple(x.clone().requires_grad_(True) for x in inputs), render


@pytest.mark.parametrize("antialias", [False, True])
def test_oracle_gradcheck(antialias):
  config = RasterConfig(tile_size=8, pixel_stride=(1, 1), antialias=antialias)
  for seed in range(12):
    inputs, render = single_tile_inputs(seed, config)
    assert torch.autograd.gradcheck(render, inputs, eps=1e-6, atol=1e-5, nondet_tol=1e-12)


def test_visibility_equals_feature_gradient():
  config = RasterConfig(compute_visibility=True, compute_point_heuristic=True)
  size = (320, 200)
  for seed in range(3):
    g, depth, feat = scene2d(seed, 1500 + 700 * seed, size, scale_factor=0.2, alpha_range=(0.2, 1.0),
                             dtype=torch.float64)
    feat.requires_grad_(True)
    raster = oracle.rasterize(g, depth, feat, size, config)
    raster.image.sum().backward()
    assert torch.allclose(feat.grad[:, 0], raster.visibility, rtol=1e-9, atol=1e-12)
    assert raster.point_heuristic.shape == (g.shape[0], 2) and (raster.point_heuristic >= 0).all()


def test_stale_tail_only_when_more_than_one_group():
  """emulate_stale_tail changes the image exactly for tiles with C > 256 and C % 256 != 0."""
  config = RasterConfig(tile_size=16)
  size = (32, 16)  # two tiles
  torch.manual_seed(0)
  n = 300
  g, depth, feat = scene2d(3, n, (16, 16), scale_factor=6.0, alpha_range=(0.02, 0.08))
  o2p = torch.cat([torch.arange(n, dtype=torch.int32), torch.arange(200, dtype=torch.int32)])
  ranges = torch.tensor([[0, n], [n, n + 200]], dtype=torch.int32)
  g2 = g.clone()
  a, wa, _ = oracle.raster_forward(g2, feat, o2p, ranges, size, config, emulate_stale_tail=True)
  b, wb, _ = oracle.raster_forward(g2, feat, o2p, ranges, size, config, emulate_stale_tail=False)
  assert not torch.equal(a[:, :16], b[:, :16]), "tile 0 (C=300) must show the stale tail"
  assert torch.equal(a[:, 16:], b[:, 16:]), "tile 1 (C=200) must not"
  # the emulated result equals blending entries [C-256, 256) a second time after the C real ones
  extra = torch.arange(n - 256, 256, dtype=torch.int32)
  o2p_manual = torch.cat([torch.arange(n, dtype=torch.int32), extra, torch.arange(200, dtype=torch.int32)])
  m = n + extra.numel()
  ranges_manual = torch.tensor([[0, m], [m, m + 200]], dtype=torch.int32)
  # walk the manual list with a tile area large enough to be a single group: tile_size 32 image 32x32 crop
  c, wc, _ = oracle.raster_forward(g2, feat, o2p_manual[:m], torch.tensor([[0, m], [0, 0]], dtype=torch.int32),
                                   size, RasterConfig(tile_size=16), emulate_stale_tail=False)
  assert rel_l2(a[:, :16], c[:, :16]) < 1e-6


def test_quantile_mode_and_weights():
  config = RasterConfig(use_alpha_blending=False, saturate_threshold=0.5)
  size = (64, 48)
  g, depth, feat = scene2d(7, 400, size, channels=1, scale_factor=2.0, alpha_range=(0.3, 0.9))
  raster = oracle.rasterize(g, depth, depth.clone(), size, config)
  assert set(np.unique(raster.image_weight.numpy()).tolist()) <= {0.0, 1.0}
  assert raster.image.min() >= 0 and raster.image.max() <= 1
