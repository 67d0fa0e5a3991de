"""View-parallel path on CPU: two `gloo` ranks, replicated gaussians, views partitioned
round-robin, ONE all-reduce of the flat gradient bucket per batch.  The per-view forward + backward is done by
the CPU oracle here (the product kernels need a GPU; tests/test_gpu_renderer.py covers them), so this checks
the host logic that bench.py and a trainer run at N > 1: partition_views, GradientBucket (autograd accumulates
straight into the flat buffer), the collective, and that the result equals the single-process sum."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taichi_gaussian_rasterizer_b200 import RasterConfig
from taichi_gaussian_rasterizer_b200.distributed import GradientBucket, all_reduce_statistics, partition_views
from util import rel_l2, scene3d

NUM_VIEWS = 5   # odd on purpose: ranks get 3 and 2 views
IMAGE_SIZE = (64, 48)


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _scene():
  from taichi_gaussian_rasterizer_b200.torch_lib.projection import join_rt, quat_to_mat
  gaussians, camera = scene3d(0, 300, image_size=IMAGE_SIZE, scale_factor=1.0, sh_degree=1)
  g = torch.Generator().manual_seed(7)
  cameras = [camera]
  for _ in range(NUM_VIEWS - 1):
    axis = torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
    angle = (torch.rand(1, generator=g) * 2 - 1) * 0.05
    q = torch.cat([axis * torch.sin(angle / 2), torch.cos(angle / 2)])
    cameras.append(camera.transformed(join_rt(quat_to_mat(q), (torch.rand(3, generator=g) * 2 - 1) * 0.05)))
  return gaussians, cameras


def _render_view_loss(gaussians, camera, config):
  """One view through the CPU oracle pipeline; returns (loss, visibility scattered to (N,))."""
  import oracle
  from oracle import torch_ref
  pts, depth, idx = torch_ref.projection_apply(*gaussians.shape_tensors(), camera.T_camera_world, camera.projection,
                                               camera.image_size, camera.depth_range, config.blur_cov,
                                               config.clamp_margin, config.alpha_threshold)
  feats = torch_ref.evaluate_sh_at(gaussians.feature, gaussians.position.detach(), idx, camera.camera_position)
  ndc = torch_ref.ndc_depth(depth, camera.near_plane, camera.far_plane)
  o2p, ranges = oracle.map_to_tiles(pts.detach().float(), ndc.detach().float(), camera.image_size, config)
  out = oracle.rasterize_with_tiles(pts, feats, o2p, ranges.view(-1, 2), camera.image_size, config)
  vis = torch.zeros(gaussians.position.shape[0], dtype=pts.dtype)
  vis[idx] = out.visibility
  return out.image.square().mean(), vis


def _params(gaussians):
  return [gaussians.position, gaussians.log_scaling, gaussians.rotation, gaussians.alpha_logit, gaussians.feature]


def _accumulate(gaussians, cameras, view_ids, config):
  gaussians.requires_grad_(True)
  bucket = GradientBucket(_params(gaussians))
  vis_total = torch.zeros(gaussians.position.shape[0])
  for v in view_ids:
    loss, vis = _render_view_loss(gaussians, cameras[v], config)
    loss.backward()
    vis_total += vis
  return bucket, vis_total


def _worker(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    torch.set_num_threads(1)
    config = RasterConfig(compute_visibility=True)
    gaussians, cameras = _scene()
    mine = partition_views(NUM_VIEWS, rank, world)
    bucket, vis = _accumulate(gaussians, cameras, mine, config)
    # autograd wrote into the flat bucket: param.grad are views of it
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in bucket.params)
    bucket.all_reduce()
    all_reduce_statistics([vis])
    torch.save({"flat": bucket.flat.clone(), "vis": vis, "views": mine}, os.path.join(out_dir, f"rank{rank}.pt"))
  finally:
    dist.destroy_process_group()


def test_partition_views_is_a_partition():
  for n in (0, 1, 5, 64):
    for world in (1, 2, 4, 8):
      parts = [partition_views(n, r, world) for r in range(world)]
      assert sorted(v for p in parts for v in p) == list(range(n))
      assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_bucket_all_reduce_is_noop_without_process_group():
  p = torch.zeros(4, 3, requires_grad=True)
  b = GradientBucket([p])
  (p.sum() * 2).backward()
  assert b.all_reduce() is None
  assert torch.equal(b.flat, torch.full((12,), 2.0)) and b.nbytes == 48


def test_symmetric_bucket_falls_back_without_nvswitch():
  """GradientBucket(symmetric=True) asks for an NVSwitch multicast mapping of the bucket (gs_multimem_all_reduce);
  without a CUDA device / NCCL group it must quietly be an ordinary bucket (reducer None), same behaviour otherwise."""
  p = torch.zeros(8, 3, requires_grad=True)
  q = torch.zeros(8, requires_grad=True)
  bucket = GradientBucket([p, q], symmetric=True)
  assert bucket.reducer is None and bucket.flat.shape == (32,)
  (p.sum() * 2 + q.sum()).backward()
  assert bucket.all_reduce() is None
  assert torch.equal(bucket.flat, torch.cat([torch.full((24,), 2.0), torch.ones(8)]))


@pytest.mark.timeout(600)
def test_view_parallel_gradients_equal_serial_sum(tmp_path):
  world = 2
  mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
  results = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
  assert results[0]["views"] == [0, 2, 4] and results[1]["views"] == [1, 3]
  # every rank holds the same reduced bucket
  assert torch.equal(results[0]["flat"], results[1]["flat"])
  assert torch.equal(results[0]["vis"], results[1]["vis"])
  # and it is the serial sum over all views (same tolerance as gradients: 1e-4 relative L2)
  config = RasterConfig(compute_visibility=True)
  gaussians, cameras = _scene()
  bucket, vis = _accumulate(gaussians, cameras, list(range(NUM_VIEWS)), config)
  assert bucket.flat.abs().sum() > 0
  assert rel_l2(results[0]["flat"], bucket.flat) < 1e-4
  assert rel_l2(results[0]["vis"], vis) < 1e-4


# ------------------------------------------------------------------------------------- reduce_early (host logic on gloo)
def _torch_flush(sink, points, staged, camera_positions, overwrite, rows=None):
  """What gs_sh_bwd_flush computes, in torch: sink (+)= sum_v staged_v (x) basis(points - camera_v)."""
  from oracle import torch_ref
  if rows is not None:
    lo, hi = rows
    sink, points, staged = sink[lo:hi], points[lo:hi], [t[lo:hi] for t in staged]
  n, k, d = sink.shape
  degree = int(round(d ** 0.5)) - 1
  total = torch.zeros_like(sink)
  for g, c in zip(staged, camera_positions):
    dirs = torch.nn.functional.normalize(points - c.reshape(1, 3), dim=1)
    total += g.unsqueeze(2) * torch_ref.rsh_cart(degree, dirs).unsqueeze(1)
  if overwrite:
    sink.copy_(total)
  else:
    sink.add_(total)


def test_deferred_flush_in_slices_equals_one_flush(monkeypatch):
  """DeferredSH.flush(chunks=, after_chunk=): the slices cover every row exactly once, in order, and give the rows of
  a single flush (GradientBucket._all_reduce_pipelined reduces a slice while the next one is formed)."""
  from taichi_gaussian_rasterizer_b200 import grad_sinks
  monkeypatch.setattr(grad_sinks, "flush_sh_views", _torch_flush)
  g = torch.Generator().manual_seed(5)
  n = 37
  points = torch.randn(n, 3, generator=g)
  views = [(torch.randn(n, 3, generator=g), torch.randn(3, generator=g)) for _ in range(3)]
  results = []
  for chunks in (1, 4):
    sink = torch.full((n, 3, 4), 7.0)
    d = grad_sinks.DeferredSH(sink)
    d.mark_clean()
    for staged, cam in views:
      d.add(staged, cam, points)
    seen = []
    assert d.flush(chunks=chunks, after_chunk=lambda lo, hi: seen.append((lo, hi)))
    assert seen[0][0] == 0 and seen[-1][1] == n and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    assert len(seen) == chunks and not d.pending
    results.append(sink)
  assert torch.equal(results[0], results[1])
  assert not grad_sinks.DeferredSH(torch.zeros(2, 3, 4)).flush(chunks=2)   # nothing pending


def _early_worker(rank, world, port, out_dir):
  from taichi_gaussian_rasterizer_b200 import grad_sinks
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    torch.set_num_threads(1)
    grad_sinks.flush_sh_views = _torch_flush   # the CUDA flush kernel's arithmetic, on the CPU
    n, views = 50, 3
    g = torch.Generator().manual_seed(3)
    points = torch.randn(n, 3, generator=g)
    geometry = torch.zeros(n, 3, requires_grad=True)
    sh = torch.zeros(n, 3, 4, requires_grad=True)
    bucket = GradientBucket([geometry, sh])
    deferred = grad_sinks.register_deferred_sh(sh, sh.grad)
    bucket.zero_()
    expected_sh = torch.zeros(n, 3, 4)
    expected_geo = torch.zeros(n, 3)
    gen = torch.Generator().manual_seed(100)
    for r in range(world):          # every rank draws every rank's inputs, and uses its own
      for v in range(views):
        staged, cam, geo = torch.randn(n, 3, generator=gen), torch.randn(3, generator=gen), torch.randn(n, 3, generator=gen)
        _torch_flush(expected_sh, points, [staged], [cam], overwrite=False)
        expected_geo += geo
        if r == rank:
          geometry.grad.add_(geo)
          deferred.add(staged, cam, points)
          if v == views - 2:
            bucket.reduce_early()   # everything but the last view: SH slice reduced now, deferred state on hold
            assert deferred.hold and not deferred.pending
    assert len(deferred.pending) == 1
    bucket.all_reduce()
    assert not deferred.hold and not deferred.pending and bucket._early is None
    grad_sinks.unregister_deferred_sh(sh)
    torch.save({"sh": sh.grad.clone(), "geo": geometry.grad.clone(), "expected_sh": expected_sh,
                "expected_geo": expected_geo}, os.path.join(out_dir, f"early{rank}.pt"))
  finally:
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_reduce_early_then_gather_equals_plain_sum(tmp_path):
  """GradientBucket.reduce_early(): SH slices all-reduced before the last view, the last view's staged colour
  gradients all-gathered and flushed by every rank — same sums as flushing everything and reducing once."""
  world = 2
  mp.spawn(_early_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
  for r in range(world):
    res = torch.load(tmp_path / f"early{r}.pt")
    assert rel_l2(res["sh"], res["expected_sh"]) < 1e-6
    assert rel_l2(res["geo"], res["expected_geo"]) < 1e-6
