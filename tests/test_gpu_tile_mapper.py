"""map_to_tiles on the GPU: overlap lists, sorted orderings and tile ranges bit-exact against the oracle
(north_star: "tile overlap lists and sorted orderings are bit-exact")."""
import pytest
import torch

import oracle
from taichi_gaussian_rasterizer_b200 import RasterConfig, map_to_tiles
from taichi_gaussian_rasterizer_b200.mapper.tile_mapper import map_to_tiles_staged
from util import rel_l2, scene2d

pytestmark = pytest.mark.gpu


def check(cuda_device, g, depth, size, cfg, depth16=False):
  o2p_ref, ranges_ref = oracle.map_to_tiles(g, depth, size, cfg, use_depth16=depth16)
  o2p, ranges = map_to_tiles(g.to(cuda_device), depth.to(cuda_device), size, cfg, use_depth16=depth16)
  assert o2p.dtype == torch.int32 and ranges.dtype == torch.int32
  assert ranges.shape == ranges_ref.shape
  assert o2p.shape == o2p_ref.shape, f"K differs: {o2p.shape} vs {o2p_ref.shape}"
  assert torch.equal(ranges.cpu(), ranges_ref)
  assert torch.equal(o2p.cpu(), o2p_ref)
  # the reference's stage order (64 bit keys, one K-sized sort) on the same kernels gives the same bits
  o2p_s, ranges_s = map_to_tiles_staged(g.to(cuda_device), depth.to(cuda_device), size, cfg, use_depth16=depth16)
  assert torch.equal(o2p_s, o2p) and torch.equal(ranges_s, ranges)
  return o2p, ranges


@pytest.mark.parametrize("seed,n,size,ts,scale", [
  (0, 1000, (320, 200), 16, 1.0),
  (1, 5000, (333, 257), 16, 3.0),      # image not a multiple of the tile
  (2, 2000, (256, 256), 8, 2.0),
  (3, 3000, (640, 480), 32, 4.0),
  (4, 50, (1024, 768), 16, 60.0),      # screen filling gaussians: warp-cooperative spans
  (5, 20000, (1024, 1024), 16, 0.5),   # BASELINE config 1 shape
])
def test_bit_exact_vs_oracle(cuda_device, seed, n, size, ts, scale):
  cfg = RasterConfig(tile_size=ts)
  g, depth, _ = scene2d(seed, n, size, scale_factor=scale)
  check(cuda_device, g, depth, size, cfg)


def test_depth16_and_ties(cuda_device):
  cfg = RasterConfig()
  size = (320, 200)
  g, depth, _ = scene2d(9, 4000, size, scale_factor=2.0)
  depth[::5] = depth[2]
  check(cuda_device, g, depth, size, cfg, depth16=False)
  check(cuda_device, g, depth, size, cfg, depth16=True)


def test_edge_cases(cuda_device):
  cfg = RasterConfig()
  size = (100, 60)
  o2p, ranges = map_to_tiles(torch.zeros((0, 7), device=cuda_device), torch.zeros((0, 1), device=cuda_device), size, cfg)
  assert o2p.shape == (0,) and ranges.shape == (4, 7, 2) and int(ranges.abs().sum()) == 0
  g = torch.tensor([[50., 30., 1., 0., 5., 3., 0.001],       # alpha below threshold: no overlaps
                    [1e6, 1e6, 1., 0., 5., 3., 0.5],          # far off screen
                    [50., 30., 0.6, 0.8, 400., 300., 0.9],    # covers everything
                    [-500., 30., 1., 0., 2., 2., 0.9],
                    [99.5, 59.5, 0., 1., 0.4, 0.2, 0.9]])     # tiny, in the last partial tile
  depth = torch.tensor([[0.5], [0.1], [0.9], [0.2], [0.0]])
  check(cuda_device, g, depth, size, cfg)
  # nothing overlaps at all
  g2 = g[:2].clone()
  o2p, ranges = check(cuda_device, g2, depth[:2], size, cfg)
  assert o2p.shape == (0,)


def test_too_many_tiles_is_rejected(cuda_device):
  cfg = RasterConfig(tile_size=8)
  with pytest.raises(AssertionError, match="exceed maximum tile count"):
    map_to_tiles(torch.rand(4, 7, device=cuda_device), torch.rand(4, 1, device=cuda_device), (4096, 2048), cfg)


def test_large_scene_properties(cuda_device):
  """Full-size property check (no oracle): ranges partition [0, K), keys sorted per tile."""
  cfg = RasterConfig()
  size = (2048, 1365)
  g, depth, _ = scene2d(21, 1_000_000, size, scale_factor=1.0)
  o2p, ranges = map_to_tiles(g.to(cuda_device), depth.to(cuda_device), size, cfg)
  r = ranges.view(-1, 2).cpu().long()
  K = o2p.shape[0]
  nonempty = r[:, 1] > r[:, 0]
  starts, ends = r[nonempty, 0], r[nonempty, 1]
  assert int(starts[0]) == 0 and int(ends[-1]) == K
  assert torch.equal(starts[1:], ends[:-1])
  d = depth.view(-1)[o2p.cpu().long()]
  # depth non-decreasing inside every tile: check all boundaries at once
  same_tile = torch.ones(K - 1, dtype=torch.bool)
  same_tile[(ends[:-1] - 1)] = False
  assert bool(((d[1:] >= d[:-1]) | ~same_tile).all())
  counts = oracle.tile_counts(g, size, cfg)
  assert int(counts.sum()) == K
  o2p_s, ranges_s = map_to_tiles_staged(g.to(cuda_device), depth.to(cuda_device), size, cfg)
  assert torch.equal(o2p_s, o2p) and torch.equal(ranges_s, ranges)


def test_fused_ndc_sort_depth_is_bit_identical(cuda_device):
  """render_projected hands LINEAR depth to the key kernel, which forms the NDC sort depth itself; the tile map must
  equal map_to_tiles on torch's ndc_depth() of the same depths bit for bit (same f32 operation sequence)."""
  from taichi_gaussian_rasterizer_b200.mapper.tile_mapper import _map_to_tiles
  from taichi_gaussian_rasterizer_b200.torch_lib.projection import inverse_ndc_depth, ndc_depth
  cfg = RasterConfig()
  size = (640, 480)
  for seed, (near, far) in enumerate([(0.1, 100.0), (0.05, 1000.0), (1.0, 7.5)]):
    g, unit, _ = scene2d(70 + seed, 200_000, size, scale_factor=1.0)
    linear = inverse_ndc_depth(unit.clamp(1e-6, 1 - 1e-6), near, far).to(cuda_device)   # depths between near and far
    gd = g.to(cuda_device)
    a_o2p, a_ranges = map_to_tiles(gd, ndc_depth(linear, near, far), size, cfg)
    b_o2p, b_ranges = _map_to_tiles(gd, linear, size, cfg, ndc_range=(near, far))
    assert torch.equal(a_ranges, b_ranges) and torch.equal(a_o2p, b_o2p)
    a16, _ = map_to_tiles(gd, ndc_depth(linear, near, far), size, cfg, use_depth16=True)
    b16, _ = _map_to_tiles(gd, linear, size, cfg, use_depth16=True, ndc_range=(near, far))
    assert torch.equal(a16, b16)


def test_capacity_bounded_mapping_matches_and_reports_overflow(cuda_device):
  """map_to_tiles(overlap_capacity=): same lists and ranges without any host read-back; beyond the capacity overlaps
  are dropped and the reported total says so."""
  cfg = RasterConfig()
  size = (200, 150)
  g, depth, _ = scene2d(3, 3000, size, scale_factor=2.0)
  gd, dd = g.to(cuda_device), depth.to(cuda_device)
  o2p, ranges = map_to_tiles(gd, dd, size, cfg)
  K = o2p.shape[0]
  total = torch.zeros(1, dtype=torch.int32, device=cuda_device)
  o2p_c, ranges_c = map_to_tiles(gd, dd, size, cfg, overlap_capacity=K + 1000, overlap_total_out=total)
  assert o2p_c.shape == (K + 1000,) and int(total.item()) == K
  assert torch.equal(o2p_c[:K], o2p) and torch.equal(ranges_c, ranges)
  o2p_s, ranges_s = map_to_tiles(gd, dd, size, cfg, overlap_capacity=K // 2, overlap_total_out=total)
  assert o2p_s.shape == (K // 2,) and int(total.item()) == K            # the caller sees K > capacity
  assert int(ranges_s.max()) <= K // 2                                  # nothing points past the buffer
  e, r = map_to_tiles(gd[:0], dd[:0], size, cfg, overlap_capacity=16, overlap_total_out=total)
  assert int(total.item()) == 0 and int(r.abs().sum()) == 0


def test_whole_step_in_a_cuda_graph(cuda_device):
  """The 2D fit step of BASELINE config 1 replayed from one CUDA graph (benchmarks/configs.py c1_graph) gives the
  gradients of the eager step."""
  import sys
  from pathlib import Path
  sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "benchmarks"))
  import configs
  step = configs.c1_graph(cuda_device)
  step()
  torch.cuda.synchronize()
  st = step.graph_state
  graph_grads = [t.grad.clone() for t in st["params"]]
  graph_loss = float(st["out"]["loss"])
  assert int(st["total"].item()) > 0
  eager = configs.c1(cuda_device)       # same seed, same scene, eager launches with the read-back of K
  eager()
  torch.cuda.synchronize()
  for a, b in zip(graph_grads, [t.grad for t in eager.params]):   # the graph computes what the eager step computes
    assert rel_l2(a, b) < 1e-5
  step()                                 # a second replay gives the same numbers again
  torch.cuda.synchronize()
  assert abs(float(st["out"]["loss"]) - graph_loss) < 1e-7
  for a, b in zip(graph_grads, [t.grad for t in st["params"]]):
    assert rel_l2(a, b) < 1e-5
