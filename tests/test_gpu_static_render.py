"""render_gaussians(..., overlap_capacity=): the path without a single host read-back (visible count and overlap
total stay on the device, point space tensors keep their capacity of N rows).  It runs the same kernels as the default
path on the same data, so the comparison is with the default path itself (which the other GPU tests hold against the
oracle): images bit-identical, gradients equal up to the order of the atomic additions; then the same step captured in
ONE CUDA graph and replayed, alone and inside GradientBucket.fused_accumulation over several views on two streams."""
import pytest
import torch

from taichi_gaussian_rasterizer_b200 import (CapturedStep, RasterConfig, evaluate_sh_views, overlap_capacity_for,
                                             render_gaussians)
from taichi_gaussian_rasterizer_b200.distributed import GradientBucket, run_views
from util import rel_l2, scene3d

pytestmark = pytest.mark.gpu

NAMES = ("position", "log_scaling", "rotation", "alpha_logit", "feature")


def _params(g):
  return [getattr(g, n) for n in NAMES]


def _scene(device, n=6000, size=(320, 240), sh=True, seed=3):
  g, cam = scene3d(seed, n, image_size=size, scale_factor=0.5, sh_degree=3 if sh else None)
  g = g.to(device=device)
  g.requires_grad_(True)
  return g, cam.to(device=device)


def _grads(g):
  return {n: getattr(g, n).grad.clone() for n in NAMES}


@pytest.mark.parametrize("use_sh,render_depth,stats", [(True, False, False), (False, False, True), (True, True, True), (False, True, False)])
def test_static_equals_default(cuda_device, use_sh, render_depth, stats):
  cfg = RasterConfig(compute_visibility=stats, compute_point_heuristic=stats)
  g, cam = _scene(cuda_device, sh=use_sh)
  torch.manual_seed(1)
  gi = None

  def run(**kw):
    nonlocal gi
    for p in _params(g):
      p.grad = None
    out = render_gaussians(g, cam, cfg, use_sh=use_sh, render_depth=render_depth, **kw)
    if gi is None:
      gi = torch.rand_like(out.image) - 0.3
    loss = (out.image * gi).sum()
    if render_depth:
      loss = loss + out.depth.mean()
    loss.backward()
    return out, _grads(g)

  ref, g_ref = run()
  V = ref.points_in_view.shape[0]
  total = torch.zeros(1, dtype=torch.int32, device=cuda_device)
  with torch.no_grad():
    from taichi_gaussian_rasterizer_b200 import map_to_tiles
    from taichi_gaussian_rasterizer_b200.torch_lib.projection import ndc_depth
    K = map_to_tiles(ref.gaussians2d, ndc_depth(ref.point_depth, cam.near_plane, cam.far_plane), cam.image_size,
                     cfg)[0].shape[0]
  out, g_out = run(overlap_capacity=K + 1000, overlap_total_out=total)
  assert int(total.item()) == K
  assert int(out.points_in_view_count.item()) == V
  assert out.points_in_view.shape[0] == g.position.shape[0]          # capacity-sized
  assert torch.equal(out.points_in_view[:V], ref.points_in_view)
  assert torch.equal(out.gaussians2d[:V].view(torch.int32), ref.gaussians2d.view(torch.int32))
  assert torch.equal(out.image, ref.image) and torch.equal(out.image_weight, ref.image_weight)
  if render_depth:
    assert torch.equal(out.depth, ref.depth)
  if stats:
    assert rel_l2(out.point_visibility[:V], ref.point_visibility) < 1e-6
    assert rel_l2(out.point_heuristic[:V], ref.point_heuristic) < 1e-5
  for n in NAMES:
    assert rel_l2(g_out[n], g_ref[n]) < 2e-6, (n, rel_l2(g_out[n], g_ref[n]))


@pytest.mark.parametrize("case", ["depth16", "median_depth", "antialias", "tile8", "tile32"])
def test_static_options_equal_default(cuda_device, case):
  """The options that change kernels or sort keys under the read-back free path: 16 bit depth keys, the quantile
  (median depth) pass, antialiased rendering, the generic kernels (tile 8 / 32).  Forward bit-identical to the default
  path, gradients to the order of the atomic additions."""
  kw, cfg_kw = {}, {}
  if case == "depth16":
    kw = dict(use_depth16=True)
  elif case == "median_depth":
    kw = dict(render_median_depth=True)
  elif case == "antialias":
    cfg_kw = dict(antialias=True, blur_cov=0.0)
  elif case == "tile8":
    cfg_kw = dict(tile_size=8, pixel_stride=(1, 1))
  elif case == "tile32":
    cfg_kw = dict(tile_size=32)
  cfg = RasterConfig(**cfg_kw)
  g, cam = _scene(cuda_device, n=4000)
  torch.manual_seed(2)
  gi = torch.rand(cam.image_size[1], cam.image_size[0], 3, device=cuda_device) - 0.3

  def run(**extra):
    for p in _params(g):
      p.grad = None
    out = render_gaussians(g, cam, cfg, use_sh=True, **kw, **extra)
    loss = (out.image * gi).sum()
    if out.median_depth is not None:
      loss = loss + out.median_depth.mean()
    loss.backward()
    return out, _grads(g)

  ref, g_ref = run()
  out, g_out = run(overlap_capacity=500_000)
  assert torch.equal(out.image, ref.image) and torch.equal(out.image_weight, ref.image_weight)
  if case == "median_depth":
    assert torch.equal(out.median_depth, ref.median_depth)
  for n in NAMES:
    assert rel_l2(g_out[n], g_ref[n]) < 2e-6, (n, rel_l2(g_out[n], g_ref[n]))


def test_static_drops_overlaps_beyond_capacity(cuda_device):
  """A capacity below K must not overrun anything: K is still reported, the image is simply missing gaussians."""
  cfg = RasterConfig()
  g, cam = _scene(cuda_device)
  total = torch.zeros(1, dtype=torch.int32, device=cuda_device)
  with torch.no_grad():
    full = render_gaussians(g, cam, cfg, use_sh=True, overlap_capacity=1 << 20, overlap_total_out=total)
    K = int(total.item())
    cut = render_gaussians(g, cam, cfg, use_sh=True, overlap_capacity=K // 2, overlap_total_out=total)
  assert int(total.item()) == K and K > 0
  assert torch.isfinite(cut.image).all()
  assert cut.image_weight.sum() < full.image_weight.sum()


def test_static_empty_view(cuda_device):
  """Nothing in view: count 0 on the device, empty tile lists, zero image, zero gradients, no read-back needed."""
  cfg = RasterConfig()
  g, cam = _scene(cuda_device, n=500)
  with torch.no_grad():
    g.alpha_logit.fill_(-20.0)   # alpha far below alpha_threshold: the projection culls every gaussian
  out = render_gaussians(g, cam, cfg, use_sh=True, overlap_capacity=4096)
  out.image.sum().backward()
  assert int(out.points_in_view_count.item()) == 0
  assert out.image.abs().max().item() == 0.0
  for n in NAMES:
    assert getattr(g, n).grad.abs().max().item() == 0.0


def test_static_step_in_one_cuda_graph(cuda_device):
  """forward + loss + backward of one view captured once, replayed with changed parameters: equals the eager default
  path on the new parameters (the graph holds no sizes that depend on the data)."""
  cfg = RasterConfig()
  g, cam = _scene(cuda_device, n=8000)
  target = torch.rand(cam.image_size[1], cam.image_size[0], 3, device=cuda_device)
  params = _params(g)
  static_grads = [torch.zeros_like(p) for p in params]
  for p, b in zip(params, static_grads):
    p.grad = b
  total = torch.zeros(1, dtype=torch.int32, device=cuda_device)
  holder = {}

  def body():
    for b in static_grads:
      b.zero_()
    out = render_gaussians(g, cam, cfg, use_sh=True, overlap_capacity=400_000, overlap_total_out=total)
    loss = torch.nn.functional.l1_loss(out.image, target)
    loss.backward()
    holder["loss"] = loss.detach()

  side = torch.cuda.Stream(device=cuda_device)
  side.wait_stream(torch.cuda.current_stream(cuda_device))
  with torch.cuda.stream(side):
    for _ in range(2):
      body()
  torch.cuda.current_stream(cuda_device).wait_stream(side)
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    body()

  with torch.no_grad():   # move the scene: V and K change, the graph does not
    g.position.add_(torch.randn_like(g.position) * 0.05)
    g.alpha_logit.add_(0.3)
  graph.replay()
  torch.cuda.synchronize()
  got = {n: b.clone() for n, b in zip(NAMES, static_grads)}
  loss_graph = holder["loss"].item()
  assert 0 < int(total.item()) <= 400_000

  for p in params:
    p.grad = None
  out = render_gaussians(g, cam, cfg, use_sh=True)
  loss = torch.nn.functional.l1_loss(out.image, target)
  loss.backward()
  assert abs(loss.item() - loss_graph) <= 1e-6 * abs(loss.item())
  for n in NAMES:
    assert rel_l2(got[n], getattr(g, n).grad) < 2e-6, (n, rel_l2(got[n], getattr(g, n).grad))


def test_static_multi_view_bucket_graph(cuda_device):
  """The benchmark's step shape: several views on two streams inside fused_accumulation (batched SH colours, deferred
  SH gradient), captured in one graph; equals the eager default path."""
  cfg = RasterConfig()
  g, cam0 = _scene(cuda_device, n=6000)
  cams = []
  for k in range(3):
    _, c = scene3d(10 + k, 10, image_size=cam0.image_size)
    cams.append(c.to(device=cuda_device))
  targets = [torch.rand(cam0.image_size[1], cam0.image_size[0], 3, device=cuda_device) for _ in cams]
  params = _params(g)
  bucket = GradientBucket(params)
  streams = [torch.cuda.Stream(device=cuda_device) for _ in range(2)]

  capacity = overlap_capacity_for(g, cams, cfg)
  assert 4096 < capacity < 10_000_000

  def step(static):
    with bucket.fused_accumulation():
      bucket.zero_()
      colors = evaluate_sh_views(g.feature, g.position, [c.camera_position for c in cams])

      def view(i):
        kw = dict(overlap_capacity=capacity) if static else {}
        out = render_gaussians(g, cams[i], cfg, use_sh=True, sh_colors=colors[i], **kw)
        loss = torch.nn.functional.l1_loss(out.image, targets[i])
        loss.backward()
        return loss.detach()
      total = run_views(len(cams), view, streams)
      bucket.flush()
    return total

  ref_loss = step(False).item()
  ref = bucket.flat.clone()

  assert not CapturedStep.capturing()
  captured = CapturedStep(lambda: step(True), device=cuda_device)   # warm-up runs on a side stream, then the capture
  torch.cuda.synchronize()
  bucket.flat.fill_(7.0)   # the replay must rebuild every gradient
  loss = captured.replay()
  torch.cuda.synchronize()
  assert rel_l2(bucket.flat, ref) < 2e-6, rel_l2(bucket.flat, ref)
  assert abs(loss.item() - ref_loss) <= 1e-6 * abs(ref_loss)
  with torch.no_grad():   # parameters change in place between replays, as an optimizer step does
    g.alpha_logit.sub_(0.2)
  loss2 = captured.replay().item()
  moved = step(False).item()
  assert abs(loss2 - moved) <= 1e-6 * abs(moved) and loss2 != ref_loss
