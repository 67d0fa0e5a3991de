#!/usr/bin/env python
"""Runs BASELINE.json's five configurations once each on one GPU and prints one JSON line per configuration
(fwd+bwd ms/frame with per-stage device times, V, K).  These are the parity-test configurations at FULL size:
they show that every one runs through the product path and what it costs; `bench.py` is the contract benchmark
(config quoted by the metric).  Usage: python benchmarks/configs.py [--only c1,c2,...] [--steps 5]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from taichi_gaussian_rasterizer_b200 import RasterConfig, _native, map_to_tiles, rasterize, render_gaussians  # noqa: E402
from taichi_gaussian_rasterizer_b200.misc.renderer2d import project_gaussians2d  # noqa: E402
from taichi_gaussian_rasterizer_b200.perspective.projection import project_to_image  # noqa: E402
from taichi_gaussian_rasterizer_b200.synthetic import baseline_scene, random_2d_gaussians  # noqa: E402
from taichi_gaussian_rasterizer_b200.torch_lib.projection import ndc_depth  # noqa: E402


def timed(fn, steps, warmup=3):
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  timer = _native.set_stage_timer(_native.StageTimer())
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(steps):
    info = fn()
  b.record()
  torch.cuda.synchronize()
  _native.set_stage_timer(None)
  stages = {k: round(v[1] / steps, 4) for k, v in sorted(timer.summary().items())}
  return a.elapsed_time(b) / steps, stages, info


def c1(dev):
  """fit_image_gaussians-style 2D fit: 20k gaussians, 1024x1024, tile 16, stats on (examples/fit_image_gaussians.py)."""
  torch.manual_seed(0)
  size = (1024, 1024)
  g = random_2d_gaussians(20_000, size, num_channels=3, scale_factor=0.5, alpha_range=(0.5, 1.0)).to(device=dev)
  g.requires_grad_(True)
  target = torch.rand(size[1], size[0], 3, device=dev)
  cfg = RasterConfig(compute_point_heuristic=True, compute_visibility=True, tile_size=16, pixel_stride=(2, 2),
                     antialias=True, blur_cov=0.0)

  def step():
    for t in (g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature):
      t.grad = None
    packed = project_gaussians2d(g)
    r = rasterize(packed, g.z_depth.clamp(0, 1), g.feature, size, cfg)
    loss = torch.nn.functional.mse_loss(torch.sigmoid(r.image), target)
    loss.backward()
    return {"V": 20_000, "K": K}
  with torch.no_grad():
    K = int(map_to_tiles(project_gaussians2d(g), g.z_depth.clamp(0, 1), size, cfg)[0].shape[0])
  step.params = (g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature)
  return step


def c1_graph(dev):
  """config 1 with the whole step (projection to 2D, tile mapping, rasterizer forward, loss, backward) replayed from ONE
  CUDA graph: rasterize(..., overlap_capacity=) keeps the overlap total on the device, so nothing synchronises."""
  torch.manual_seed(0)
  size = (1024, 1024)
  g = random_2d_gaussians(20_000, size, num_channels=3, scale_factor=0.5, alpha_range=(0.5, 1.0)).to(device=dev)
  g.requires_grad_(True)
  params = (g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature)
  target = torch.rand(size[1], size[0], 3, device=dev)
  cfg = RasterConfig(compute_point_heuristic=True, compute_visibility=True, tile_size=16, pixel_stride=(2, 2),
                     antialias=True, blur_cov=0.0)
  with torch.no_grad():
    K = int(map_to_tiles(project_gaussians2d(g), g.z_depth.clamp(0, 1), size, cfg)[0].shape[0])
  capacity = int(K * 1.5)
  total = torch.zeros(1, dtype=torch.int32, device=dev)
  out = {}

  def body():
    packed = project_gaussians2d(g)
    r = rasterize(packed, g.z_depth.clamp(0, 1), g.feature, size, cfg, overlap_capacity=capacity,
                  overlap_total_out=total)
    loss = torch.nn.functional.mse_loss(torch.sigmoid(r.image), target)
    loss.backward()
    out["loss"] = loss.detach()

  side = torch.cuda.Stream(device=dev)
  side.wait_stream(torch.cuda.current_stream(dev))
  with torch.cuda.stream(side):   # warm-up on a side stream, as graph capture requires
    for _ in range(3):
      for t in params:
        t.grad = None
      body()
  torch.cuda.current_stream(dev).wait_stream(side)
  for t in params:
    t.grad = None
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    body()

  def step():
    graph.replay()
    return {"V": 20_000, "K": K, "capacity": capacity}
  # everything the graph reads must outlive it: the replay dereferences the captured addresses
  step.graph_state = dict(params=params, out=out, total=total, keep_alive=(g, target, graph, body))
  return step


def scene3d(name, dev):
  """BASELINE.json configuration `name` (synthetic.baseline_scene: shared with bench.py and the full-size parity tests)."""
  g, cam, spec = baseline_scene(name)
  g = g.to(device=dev)
  g.requires_grad_(True)
  return g, cam.to(device=dev)


def count_overlaps(g, cam, cfg):
  """(V, K) of the view: visible gaussians and tile overlaps (recorded next to the timings)."""
  with torch.no_grad():
    g2d, depth, idx = project_to_image(g, cam, cfg)
    o2p, _ = map_to_tiles(g2d, ndc_depth(depth, cam.near_plane, cam.far_plane), cam.image_size, cfg)
  return int(idx.shape[0]), int(o2p.shape[0])


def render_step(g, cam, cfg, **kw):
  V, K = count_overlaps(g, cam, cfg)

  def step():
    for t in (g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature):
      t.grad = None
    r = render_gaussians(g, cam, cfg, **kw)
    loss = r.image.abs().mean()
    if r.depth is not None:
      loss = loss + r.depth.mean() * 1e-3
    loss.backward()
    return {"V": V, "K": K}
  return step


def render_step_graph(g, cam, cfg, **kw):
  """render_step with the view going through render_gaussians(..., overlap_capacity=) — nothing is read back — and the
  whole step (forward, loss, backward) replayed from ONE CUDA graph.  The gradients are assigned, not accumulated: they
  are None when the graph is captured, so every replay rewrites the tensors the capture allocated."""
  from taichi_gaussian_rasterizer_b200 import CapturedStep, overlap_capacity_for
  V, K = count_overlaps(g, cam, cfg)
  capacity = overlap_capacity_for(g, [cam], cfg)
  params = (g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature)
  total = torch.zeros(1, dtype=torch.int32, device=g.position.device)

  def body():
    r = render_gaussians(g, cam, cfg, overlap_capacity=capacity, overlap_total_out=total, **kw)
    loss = r.image.abs().mean()
    if r.depth is not None:
      loss = loss + r.depth.mean() * 1e-3
    loss.backward()

  def no_grads():
    for t in params:
      t.grad = None
  captured = CapturedStep(body, device=g.position.device, before_capture=no_grads)

  def step():
    captured.replay()
    return {"V": V, "K": K, "capacity": capacity}
  step.graph_state = dict(params=params, total=total, keep_alive=(g, cam, captured))
  return step


def c2_graph(dev):
  """config 2 as one CUDA graph per frame (read-back free render_gaussians)."""
  g, cam = scene3d("c2", dev)
  return render_step_graph(g, cam, RasterConfig(), use_sh=True)


def c3_graph(dev):
  """config 3 (6M gaussians, visibility + split/prune stats) as one CUDA graph per frame (read-back free render_gaussians)."""
  g, cam = scene3d("c3", dev)
  return render_step_graph(g, cam, RasterConfig(compute_visibility=True, compute_point_heuristic=True), use_sh=True)


def c4_graph(dev):
  """config 4 (34 channels, 4K) as one CUDA graph per frame (read-back free render_gaussians)."""
  g, cam = scene3d("c4", dev)
  return render_step_graph(g, cam, RasterConfig(), use_sh=False, render_depth=True)


def c5_graph(dev):
  """one view of config 5 as one CUDA graph per frame (read-back free render_gaussians)."""
  g, cam = scene3d("c5", dev)
  return render_step_graph(g, cam, RasterConfig(), use_sh=True)


def c2(dev):
  """render_gaussians 3D: 1M gaussians, SH degree 3, 1920x1080."""
  g, cam = scene3d("c2", dev)
  return render_step(g, cam, RasterConfig(), use_sh=True)


def c3(dev):
  """bicycle-scale: 6M gaussians, log-normal scales / bimodal opacity, 2048x1365, visibility + split/prune stats."""
  g, cam = scene3d("c3", dev)
  return render_step(g, cam, RasterConfig(compute_visibility=True, compute_point_heuristic=True), use_sh=True)


def c4(dev):
  """feature lifting: 2M gaussians, 32-channel features + depth / depth variance, 3840x2160."""
  g, cam = scene3d("c4", dev)
  return render_step(g, cam, RasterConfig(), use_sh=False, render_depth=True)


def c5(dev):
  """one rank's share of the batched multi-view step: 8 of 64 cameras x 3M gaussians, 1600x1064 (see bench.py --gpus)."""
  g, cam = scene3d("c5", dev)
  return render_step(g, cam, RasterConfig(), use_sh=True)


CONFIGS = {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}
# not part of the default list: python benchmarks/configs.py --only c1,c1_graph,c3,c3_graph
EXTRA = {"c1_graph": c1_graph, "c2_graph": c2_graph, "c3_graph": c3_graph, "c4_graph": c4_graph, "c5_graph": c5_graph}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--only", default=",".join(CONFIGS))
  ap.add_argument("--steps", type=int, default=5)
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  for name in args.only.split(","):
    try:
      maker = CONFIGS.get(name) or EXTRA[name]
      step = maker(dev)
      ms, stages, info = timed(step, args.steps)
      print(json.dumps({"config": name, "what": " ".join(maker.__doc__.split()), "ms_per_frame_fwd_bwd": round(ms, 3),
                        "stage_ms": stages, **info, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}),
            flush=True)
    except Exception as e:   # noqa: BLE001 - report and continue with the next configuration
      print(json.dumps({"config": name, "error": f"{type(e).__name__}: {e}"}), flush=True)
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
  main()
