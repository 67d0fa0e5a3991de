"""Oracle checks for the tile mapper (CPU).  The reference's tests never pin tile maps, so
the oracle is checked against an independent brute-force numpy restatement of the OBB predicate over
ALL tiles (taichi_lib/grid_query.py:9-91), the ordering contract (tile, depth bits, gaussian index) via
numpy's stable sort, and the host image of the CUDA predicate bit for bit."""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from taichi_gaussian_rasterizer_b200 import RasterConfig, _native, pad_to_tile
from util import scene2d


def brute_force_overlaps(g, image_size, config):
  """float64 numpy: every tile of the padded image against every gaussian; returns a bool (N, TH, TW)."""
  ts = config.tile_size
  w, h = pad_to_tile(image_size, ts)
  tw, th = w // ts, h // ts
  g = g.double().numpy()
  mean, axis, sigma, alpha = g[:, 0:2], g[:, 2:4], g[:, 4:6], g[:, 6]
  gs = np.sqrt(2 * np.log(alpha / config.alpha_threshold))
  scale = sigma * gs[:, None]
  axis2 = np.stack([-axis[:, 1], axis[:, 0]], -1)
  ext = np.sqrt((axis * scale[:, 0:1]) ** 2 + (axis2 * scale[:, 1:2]) ** 2)
  lo, hi = mean - ext, mean + ext
  min_t = np.maximum(np.floor(lo / ts), 0)
  max_t = np.minimum(np.maximum(np.ceil(hi / ts), min_t + 1), np.array([tw, th]))
  out = np.zeros((g.shape[0], th, tw), dtype=bool)
  ys, xs = np.meshgrid(np.arange(th), np.arange(tw), indexing='ij')
  for i in range(g.shape[0]):
    if not np.isfinite(gs[i]):
      continue
    in_span = (xs >= min_t[i, 0]) & (xs < max_t[i, 0]) & (ys >= min_t[i, 1]) & (ys < max_t[i, 1])
    corners_x = np.stack([xs, xs + 1, xs + 1, xs], -1) * ts - mean[i, 0]
    corners_y = np.stack([ys, ys, ys + 1, ys + 1], -1) * ts - mean[i, 1]
    sep = np.zeros_like(in_span)
    for ax_v, sc in ((axis[i], scale[i, 0]), (axis2[i], scale[i, 1])):
      loc = (corners_x * ax_v[0] + corners_y * ax_v[1]) / sc
      sep |= (loc.min(-1) > 1) | (loc.max(-1) < -1)
    out[i] = in_span & ~sep
  return out


@pytest.mark.parametrize("seed,n,size,ts,scale", [(0, 300, (200, 120), 16, 1.0), (1, 500, (333, 257), 16, 3.0),
                                                  (2, 200, (128, 128), 8, 2.0), (3, 200, (400, 300), 32, 4.0)])
def test_counts_match_brute_force(seed, n, size, ts, scale):
  cfg = RasterConfig(tile_size=ts)
  g, depth, _ = scene2d(seed, n, size, scale_factor=scale)
  counts = oracle.tile_counts(g, size, cfg)
  brute = brute_force_overlaps(g, size, cfg)
  bf = torch.from_numpy(brute.reshape(n, -1).sum(1)).int()
  # float32 predicate vs float64 brute force: borderline tiles may differ on a handful of gaussians
  assert (counts != bf).float().mean() < 0.01
  assert abs(int(counts.sum()) - int(bf.sum())) <= max(2, 0.002 * int(bf.sum()))


@pytest.mark.parametrize("depth16", [False, True])
def test_ordering_contract(depth16):
  cfg = RasterConfig()
  size = (320, 200)
  g, depth, _ = scene2d(5, 2000, size, scale_factor=2.0)
  depth[::7] = depth[3]   # force exact depth ties so stability matters
  o2p, ranges = oracle.map_to_tiles(g, depth, size, cfg, use_depth16=depth16)
  counts = oracle.tile_counts(g, size, cfg)
  cum, total = oracle.full_cumsum(counts)
  keys, values = oracle.tile_emit_keys(g, depth, cum[:-1], total, size, cfg, depth16)
  k = keys.numpy().astype(np.uint64) & np.uint64((1 << (32 if depth16 else 48)) - 1)
  order = np.argsort(k, kind='stable')
  assert (values.numpy()[order] == o2p.numpy()).all()
  # ranges partition [0, K) by tile, tiles ascending, empty tiles [0, 0]
  shift = 16 if depth16 else 32
  tiles = (k[order] >> np.uint64(shift)).astype(np.int64)
  r = ranges.view(-1, 2).numpy()
  for t in range(r.shape[0]):
    idx = np.nonzero(tiles == t)[0]
    if len(idx) == 0:
      assert tuple(r[t]) == (0, 0)
    else:
      assert tuple(r[t]) == (idx[0], idx[-1] + 1)
  # within a tile: depth ascending, ties by gaussian index
  d = depth.view(-1).numpy()
  for t in np.unique(tiles)[:50]:
    s, e = r[t]
    ids = o2p.numpy()[s:e]
    dk = d[ids] if not depth16 else (np.clip(d[ids], 0, 1) * np.float32(65535.0)).astype(np.uint32)
    pairs = list(zip(dk.tolist(), ids.tolist()))
    assert pairs == sorted(pairs)


def test_edge_cases():
  cfg = RasterConfig()
  size = (100, 60)  # not a multiple of the tile size
  # empty input
  o2p, ranges = oracle.map_to_tiles(torch.zeros((0, 7)), torch.zeros((0, 1)), size, cfg)
  assert o2p.shape == (0,) and ranges.shape == (4, 7, 2) and int(ranges.abs().sum()) == 0
  # alpha below threshold (NaN scale) -> defined as zero overlaps; far off screen -> none or tile-clamped
  g = torch.tensor([[50., 30., 1., 0., 5., 3., 0.001],
                    [1e6, 1e6, 1., 0., 5., 3., 0.5],
                    [50., 30., 0.6, 0.8, 400., 300., 0.9],      # covers the whole image
                    [-500., 30., 1., 0., 2., 2., 0.9]])
  c = oracle.tile_counts(g, size, cfg)
  assert c[0] == 0 and c[1] == 0
  assert c[2] == 4 * 7
  assert c[3] <= 1


def test_cuda_host_image_tile_query_is_identical_to_oracle():
  lib = _native.lib()
  fn = lib.gs_selftest_tile_query
  fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
  fn.restype = ctypes.c_int
  cfg = RasterConfig()
  size = (333, 257)
  padded = pad_to_tile(size, cfg.tile_size)
  g, depth, _ = scene2d(11, 1500, size, scale_factor=3.0)
  counts = oracle.tile_counts(g, size, cfg)
  cum, total = oracle.full_cumsum(counts)
  keys, values = oracle.tile_emit_keys(g, depth, cum[:-1], total, size, cfg, False)
  tile_ids = (keys.numpy().astype(np.uint64) >> np.uint64(32)).astype(np.int32)
  buf = torch.zeros(4096, dtype=torch.int32)
  for i in range(g.shape[0]):
    n = fn(g[i].data_ptr(), padded[0], padded[1], cfg.tile_size, ctypes.c_float(cfg.alpha_threshold),
           buf.data_ptr(), 4096)
    assert n == int(counts[i])
    assert (buf[:n].numpy() == tile_ids[int(cum[i]):int(cum[i]) + n]).all()
