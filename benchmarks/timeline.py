#!/usr/bin/env python
"""GPU timeline of the bench workload: per kernel device time, share of the frame, and how much of the frame the
GPU sits idle (host syncs, launch gaps).  Uses torch.profiler (CUPTI), so the absolute times carry tracing
overhead: read the SHARES and the idle fraction; bench.py is the timing authority.

  python benchmarks/timeline.py [--views 4] [--num-gaussians 3000000]
"""
import argparse
import collections
import json
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from taichi_gaussian_rasterizer_b200 import RasterConfig, evaluate_sh_views, render_gaussians  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--views", type=int, default=4)
  ap.add_argument("--num-gaussians", type=int, default=0)
  args = ap.parse_args()
  from taichi_gaussian_rasterizer_b200.synthetic import BASELINE_SCENES
  W = dict(bench.WORKLOADS["bench"], image_size=BASELINE_SCENES["bench"]["image_size"])
  dev = torch.device("cuda:0")
  g_cpu, cams, _ = bench.build_scene(W["scene"], W["seed"], args.views, args.num_gaussians or None)
  g = g_cpu.to(device=dev)
  g.requires_grad_(True)
  cams = [c.to(device=dev) for c in cams]
  w, h = W["image_size"]
  targets = [torch.rand(h, w, 3, device=dev) for _ in cams]
  cfg = RasterConfig(tile_size=W["tile_size"])

  # as bench.py: gradients of all views accumulate in one flat bucket, inside the kernels
  from taichi_gaussian_rasterizer_b200.distributed import GradientBucket
  bucket = GradientBucket([g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature])

  def step():
    with bucket.fused_accumulation():
      bucket.zero_()
      colors = evaluate_sh_views(g.feature, g.position, [c.camera_position for c in cams])
      for cam, tgt, col in zip(cams, targets, colors):
        r = render_gaussians(g, cam, cfg, use_sh=True, sh_colors=col)
        torch.nn.functional.l1_loss(r.image, tgt).backward()

  for _ in range(3):
    step()
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
  evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
  evs.sort(key=lambda e: e.time_range.start)
  t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
  busy, cur_end = 0.0, t0
  for e in evs:   # union of kernel intervals
    s, en = max(e.time_range.start, cur_end), e.time_range.end
    if en > s:
      busy += en - s
      cur_end = en
  gaps, cur_end, prev = [], t0, None
  for e in evs:
    if prev is not None and e.time_range.start > cur_end:
      gaps.append((e.time_range.start - cur_end, prev.name[:60], e.name[:60]))
    if e.time_range.end >= cur_end:
      cur_end, prev = e.time_range.end, e
  gap_by_pair = collections.defaultdict(lambda: [0, 0.0])
  for g_us, a, b in gaps:
    gap_by_pair[(a, b)][0] += 1
    gap_by_pair[(a, b)][1] += g_us
  agg = collections.OrderedDict()
  for e in evs:
    a = agg.setdefault(e.name, [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
  span = t1 - t0
  rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
  ours = sum(v[1] for k, v in rows if k.startswith("void gs::") or k.startswith("gs::"))
  print(json.dumps({"frames": args.views, "span_ms_per_frame": span / 1e3 / args.views,
                    "gpu_busy_ms_per_frame": busy / 1e3 / args.views, "idle_fraction": 1 - busy / span,
                    "libgsplat_kernels_ms_per_frame": ours / 1e3 / args.views,
                    "other_kernels_ms_per_frame": (sum(v[1] for _, v in rows) - ours) / 1e3 / args.views}))
  print("largest idle gaps (us per frame, count per frame, kernel before -> kernel after):")
  for (a, b), (n, us) in sorted(gap_by_pair.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {us / args.views:7.1f} us  x{n / args.views:4.1f}  {a}  ->  {b}")
  for name, (n, us) in rows[:12]:
    print(f"{us / 1e3 / args.views:8.4f} ms/frame  {n / args.views:6.1f} launches/frame  {100 * us / span:5.1f}%  {name[:110]}")


if __name__ == "__main__":
  main()
