#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/*.npz FROM THE REFERENCE ITSELF (needs /root/reference).

  python tests/golden/make_golden.py [--only tile_map,raster,raster2,projection,sh,optim,morton] [--quick]

Two sources, both the reference's own code, unmodified, imported from /root/reference:

* the pure-torch implementation the reference's tests use as THEIR oracle
  (taichi_splatting/torch_lib/projection.py ``apply``, torch_lib/spherical_harmonics.py ``evaluate_sh_at``):
  outputs and autograd gradients in float64 -> projection.npz, sh.npz;
* the reference's Taichi kernels executed by the emulator in ti_emu.py (Taichi itself is not installable
  here): ``map_to_tiles`` (tile_overlaps / generate_sort_keys / find_ranges kernels + the OBB grid query),
  ``rasterize_with_tiles`` forward and backward (rasterizer/forward.py, backward.py, incl. shared-memory
  staging, warp votes and shuffle reductions), ``project_kernel`` and ``evaluate_sh_at_kernel`` forward
  -> tile_map.npz, raster.npz and the ``ti_*`` entries of projection.npz / sh.npz; and the reference's optimizer
  classes (optim/fractional.py, optim/visibility_aware.py with their Taichi step kernels) -> optim.npz.

Inputs are produced by the reference's own generators (taichi_splatting/tests/random_data.py) with the seeds
listed below and stored in the fixtures, so the tests in tests/test_golden.py need neither /root/reference nor
the emulator.  Nothing from oracle/ or from the product package is used to make these files.
"""
import argparse
import os
import sys
import time
from pathlib import Path

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")   # the reference's @torch.compile helpers run eagerly (no Inductor here)

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import ti_emu  # noqa: E402

ti_emu.install()

from taichi_splatting.data_types import RasterConfig  # noqa: E402
from taichi_splatting.mapper.tile_mapper import map_to_tiles  # noqa: E402
from taichi_splatting.rasterizer.function import rasterize_with_tiles  # noqa: E402
from taichi_splatting.misc.renderer2d import project_gaussians2d  # noqa: E402
import taichi_splatting.perspective.projection as ti_proj  # noqa: E402
import taichi_splatting.torch_lib.projection as torch_proj  # noqa: E402
import taichi_splatting.torch_lib.spherical_harmonics as torch_sh  # noqa: E402
import taichi_splatting.spherical_harmonics as ti_sh  # noqa: E402
from taichi_splatting.tests.random_data import random_2d_gaussians, random_3d_gaussians, random_camera  # noqa: E402


def np_(t):
  return t.detach().cpu().numpy().copy()   # a copy: in-place optimizer updates must not alias stored snapshots


def grad_(t):
  return np_(t.grad if t.grad is not None else torch.zeros_like(t))   # no dependence (e.g. SH degree 0) = zero


def scene2d(seed, n, image_size, channels, scale_factor, alpha_range=(0.1, 0.9)):
  """The reference's 2D generator (tests/random_data.py:80-105) packed as rasterize() expects
  (misc/renderer2d.py:17-33, examples/fit_image_gaussians.py:108-112)."""
  torch.manual_seed(seed)
  g = random_2d_gaussians(n, image_size, num_channels=channels, scale_factor=scale_factor, alpha_range=alpha_range)
  packed = project_gaussians2d(g).to(torch.float32).contiguous()
  return packed, g.z_depth.clamp(0, 1).to(torch.float32).contiguous(), g.feature.to(torch.float32).contiguous()


# ----------------------------------------------------------------------------------------------- tile map
TILE_CASES = [  # seed, n, (w, h), tile_size, scale_factor, use_depth16
  (0, 200, (70, 45), 8, 1.5, False),
  (1, 300, (128, 96), 16, 1.5, False),
  (2, 50, (33, 31), 16, 1.0, False),
  (3, 400, (160, 100), 16, 3.0, False),
  (4, 150, (64, 64), 32, 2.0, False),
  (5, 250, (96, 64), 16, 1.5, True),
  (6, 120, (40, 24), 8, 0.5, True),
]


def make_tile_map(out):
  for i, (seed, n, size, ts, sf, d16) in enumerate(TILE_CASES):
    g, depth, _ = scene2d(seed, n, size, 3, sf)
    t0 = time.time()
    o2p, ranges = map_to_tiles(g, depth, size, RasterConfig(tile_size=ts), use_depth16=d16)
    print(f"tile_map case {i}: n={n} {size} tile {ts} depth16={d16} K={o2p.shape[0]} ({time.time() - t0:.1f}s)")
    out.update({f"c{i}_gaussians": np_(g), f"c{i}_depth": np_(depth), f"c{i}_image_size": np.array(size),
                f"c{i}_tile_size": np.array(ts), f"c{i}_use_depth16": np.array(int(d16)),
                f"c{i}_overlap_to_point": np_(o2p), f"c{i}_tile_ranges": np_(ranges)})
  out["num_cases"] = np.array(len(TILE_CASES))


# ----------------------------------------------------------------------------------------------- rasterizer
RASTER_CASES = [  # seed, n, (w, h), tile_size, pixel_stride, F, antialias, scale_factor
  (0, 30, (16, 16), 8, (1, 1), 3, False, 1.5),
  (1, 120, (16, 8), 8, (1, 1), 3, False, 1.5),     # tile lists longer than one block: stale shared slots (Q1)
  (2, 60, (32, 32), 16, (2, 2), 3, False, 3.0),    # default tile / stride, 4 tiles
  (3, 40, (16, 16), 8, (1, 1), 2, True, 1.5),      # antialiased pdf
  (4, 90, (24, 13), 8, (1, 1), 1, False, 2.0),     # image not a multiple of the tile, F = 1
  (5, 48, (16, 16), 16, (2, 2), 5, False, 4.0),    # F = 5 (depth + depth^2 + rgb)
]
QUICK_RASTER = [0, 3]


def make_raster(out, quick=False):
  cases = [c for i, c in enumerate(RASTER_CASES) if not quick or i in QUICK_RASTER]
  for i, (seed, n, size, ts, stride, F, aa, sf) in enumerate(cases):
    g, depth, feat = scene2d(seed, n, size, F, sf)
    cfg = RasterConfig(tile_size=ts, pixel_stride=stride, antialias=aa, compute_visibility=True,
                       compute_point_heuristic=True)
    o2p, ranges = map_to_tiles(g, depth, size, cfg)
    torch.manual_seed(1000 + seed)
    grad_image = torch.rand(size[1], size[0], F)
    gg, ff = g.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    t0 = time.time()
    r = rasterize_with_tiles(gg, ff, o2p, ranges.view(-1, 2), size, cfg)
    t1 = time.time()
    (r.image * grad_image).sum().backward()
    counts = ranges[..., 1] - ranges[..., 0]
    print(f"raster case {i}: n={n} {size} tile {ts} stride {stride} F={F} aa={aa} max tile list {int(counts.max())} "
          f"(fwd {t1 - t0:.0f}s bwd {time.time() - t1:.0f}s)")
    out.update({f"c{i}_gaussians": np_(g), f"c{i}_features": np_(feat), f"c{i}_image_size": np.array(size),
                f"c{i}_tile_size": np.array(ts), f"c{i}_pixel_stride": np.array(stride), f"c{i}_antialias": np.array(int(aa)),
                f"c{i}_overlap_to_point": np_(o2p), f"c{i}_tile_ranges": np_(ranges), f"c{i}_grad_image": np_(grad_image),
                f"c{i}_image": np_(r.image), f"c{i}_image_weight": np_(r.image_weight),
                f"c{i}_visibility": np_(r.visibility), f"c{i}_point_heuristic": np_(r.point_heuristic),
                f"c{i}_grad_gaussians": np_(gg.grad), f"c{i}_grad_features": np_(ff.grad)})
  out["num_cases"] = np.array(len(cases))


# Round 2: the configuration the measured kernels run (tile 16, stride (2, 2), statistics on) with tile lists LONGER
# than one group of 256 and C mod 256 != 0 (the reference's stale shared-memory slots, SURVEY Q1, now at tile 16 —
# case 1 above only reaches it at tile 8), three groups per tile, 34 feature channels (BASELINE config 4) and the
# antialiased pdf at tile 16.  Low opacities keep the pixels unsaturated so that the re-read stale slots show in the
# image.  The emulator needs 5-30 minutes per case; `--cases` selects some, `--merge` joins per case files.
RASTER2_CASES = [  # seed, n, (w, h), tile_size, pixel_stride, F, antialias, scale_factor, alpha_range
  (6, 330, (16, 16), 16, (2, 2), 3, False, 6.0, (0.02, 0.1)),     # one tile, C = 330: two groups, 182 stale slots
  (7, 48, (16, 16), 16, (2, 2), 34, False, 4.0, (0.1, 0.9)),      # F = 34
  (8, 620, (32, 16), 16, (2, 2), 3, False, 6.0, (0.02, 0.08)),    # two tiles, three groups each
  (9, 300, (16, 16), 16, (2, 2), 3, True, 6.0, (0.02, 0.1)),      # antialiased, two groups
]


def make_raster2(out, only=None):
  for i, (seed, n, size, ts, stride, F, aa, sf, arange) in enumerate(RASTER2_CASES):
    if only is not None and i not in only:
      continue
    g, depth, feat = scene2d(seed, n, size, F, sf, alpha_range=arange)
    cfg = RasterConfig(tile_size=ts, pixel_stride=stride, antialias=aa, compute_visibility=True,
                       compute_point_heuristic=True, blur_cov=0.0 if aa else 0.3)
    o2p, ranges = map_to_tiles(g, depth, size, cfg)
    torch.manual_seed(1000 + seed)
    grad_image = torch.rand(size[1], size[0], F)
    gg, ff = g.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    t0 = time.time()
    r = rasterize_with_tiles(gg, ff, o2p, ranges.view(-1, 2), size, cfg)
    t1 = time.time()
    (r.image * grad_image).sum().backward()
    counts = (ranges[..., 1] - ranges[..., 0]).view(-1)
    print(f"raster2 case {i}: n={n} {size} tile {ts} stride {stride} F={F} aa={aa} tile lists {counts.tolist()} "
          f"(fwd {t1 - t0:.0f}s bwd {time.time() - t1:.0f}s)", flush=True)
    out.update({f"c{i}_gaussians": np_(g), f"c{i}_features": np_(feat), f"c{i}_image_size": np.array(size),
                f"c{i}_tile_size": np.array(ts), f"c{i}_pixel_stride": np.array(stride), f"c{i}_antialias": np.array(int(aa)),
                f"c{i}_overlap_to_point": np_(o2p), f"c{i}_tile_ranges": np_(ranges), f"c{i}_grad_image": np_(grad_image),
                f"c{i}_image": np_(r.image), f"c{i}_image_weight": np_(r.image_weight),
                f"c{i}_visibility": np_(r.visibility), f"c{i}_point_heuristic": np_(r.point_heuristic),
                f"c{i}_grad_gaussians": np_(gg.grad), f"c{i}_grad_features": np_(ff.grad)})
  out["num_cases"] = np.array(len(RASTER2_CASES))


# ----------------------------------------------------------------------------------------------- projection
PROJ_CASES = [  # seed, n, blur_cov  (generator arguments of tests/test_projection.py:24-34)
  (0, 400, 0.0), (1, 700, 0.3), (2, 150, 0.3), (3, 1000, 0.0),
]


def make_projection(out):
  for i, (seed, n, blur) in enumerate(PROJ_CASES):
    torch.manual_seed(seed)
    camera = random_camera()
    gaussians = random_3d_gaussians(n=n, camera_params=camera, margin=0.5, scale_factor=0.1)
    tensors32 = [gaussians.position, gaussians.log_scaling, gaussians.rotation, gaussians.alpha_logit,
                 camera.T_camera_world, camera.projection]
    tensors32 = [t.to(torch.float32).contiguous() for t in tensors32]
    size, drange = tuple(int(x) for x in camera.image_size), tuple(float(x) for x in camera.depth_range)

    # (a) the reference's torch implementation, float64, with autograd gradients
    t64 = [t.to(torch.float64).clone().requires_grad_(True) for t in tensors32]
    pts, depth, idx = torch_proj.apply(*t64, size, drange, blur_cov=blur)
    torch.manual_seed(2000 + seed)
    gp, gd = torch.randn_like(pts), torch.randn_like(depth)
    ((pts * gp).sum() + (depth * gd).sum()).backward()
    # (b) the reference's Taichi project_kernel (f32 and f64) under the emulator: forward only
    ti32 = ti_proj.apply(*tensors32, size, drange, blur_cov=blur)
    ti64 = ti_proj.apply(*[t.detach() for t in t64], size, drange, blur_cov=blur)
    print(f"projection case {i}: n={n} blur={blur} visible torch {idx.shape[0]} taichi f32 {ti32[2].shape[0]} "
          f"f64 {ti64[2].shape[0]}; max |taichi f64 - torch f64| = {(ti64[0] - pts).abs().max().item():.2e}")
    names = ["position", "log_scaling", "rotation", "alpha_logit", "T_camera_world", "projection"]
    for nm, t, t6 in zip(names, tensors32, t64):
      out[f"c{i}_{nm}"] = np_(t)
      out[f"c{i}_grad_{nm}"] = grad_(t6)
    out.update({f"c{i}_image_size": np.array(size), f"c{i}_depth_range": np.array(drange), f"c{i}_blur_cov": np.array(blur),
                f"c{i}_torch_points": np_(pts), f"c{i}_torch_depth": np_(depth), f"c{i}_torch_indexes": np_(idx),
                f"c{i}_grad_out_points": np_(gp), f"c{i}_grad_out_depth": np_(gd),
                f"c{i}_ti32_points": np_(ti32[0]), f"c{i}_ti32_depth": np_(ti32[1]), f"c{i}_ti32_indexes": np_(ti32[2]),
                f"c{i}_ti64_points": np_(ti64[0]), f"c{i}_ti64_depth": np_(ti64[1]), f"c{i}_ti64_indexes": np_(ti64[2])})
  out["num_cases"] = np.array(len(PROJ_CASES))


# ----------------------------------------------------------------------------------------------- SH
SH_CASES = [(0, 60, 3, 3), (1, 101, 3, 2), (2, 40, 1, 1), (3, 80, 2, 0), (4, 90, 3, 3)]  # seed, n, K, degree


def make_sh(out):
  for i, (seed, n, k, deg) in enumerate(SH_CASES):
    torch.manual_seed(seed)
    params = torch.rand(n, k, (deg + 1) ** 2)
    points = torch.randn(n, 3)
    cam = torch.randn(3)
    idx = torch.randint(0, n, (max(n // 2, 1),))      # repeated indexes, as tests/test_spherical_harmonics.py
    t64 = [t.to(torch.float64).clone().requires_grad_(True) for t in (params, points, cam)]
    o = torch_sh.evaluate_sh_at(t64[0], t64[1], idx, t64[2])
    torch.manual_seed(3000 + seed)
    go = torch.randn_like(o)
    (o * go).sum().backward()
    ti32 = ti_sh.sh_function(deg, k, torch.float32).apply(params, points, idx, cam)
    print(f"sh case {i}: n={n} K={k} degree={deg}; max |taichi f32 - torch f64| = "
          f"{(ti32.double() - o).abs().max().item():.2e}")
    out.update({f"c{i}_params": np_(params), f"c{i}_points": np_(points), f"c{i}_camera_pos": np_(cam),
                f"c{i}_indexes": np_(idx), f"c{i}_torch_out": np_(o), f"c{i}_grad_out": np_(go),
                f"c{i}_grad_params": grad_(t64[0]), f"c{i}_grad_points": grad_(t64[1]),
                f"c{i}_grad_camera_pos": grad_(t64[2]), f"c{i}_ti32_out": np_(ti32)})
  out["num_cases"] = np.array(len(SH_CASES))


# ----------------------------------------------------------------------------------------------- optimizers
OPTIM_CLASSES = ["FractionalAdam", "FractionalLaProp", "SparseAdam", "SparseLaProp", "VisibilityAwareAdam",
                 "VisibilityAwareLaProp"]
OPTIM_STEPS, OPTIM_N = 4, 60


def optim_problem(seed):
  """Four parameter groups shaped like the 2D fit of examples/fit_image_gaussians.py:262-273 (position in the local
  basis, scalar log-scales with a per column mask, vector features with per point rates, an (N, 3, 4) scalar block)."""
  torch.manual_seed(seed)
  n = OPTIM_N
  params = dict(position=torch.randn(n, 2), log_scaling=torch.randn(n, 2), feature=torch.randn(n, 3),
                sh=torch.randn(n, 3, 4))
  groups = dict(position=dict(lr=0.1, type="local_vector"),
                log_scaling=dict(lr=0.05, type="scalar", mask_lr=torch.tensor([1.0, 0.25])),
                feature=dict(lr=0.02, type="vector", point_lr=torch.rand(n) + 0.5),
                sh=dict(lr=0.01, type="scalar"))
  steps = []
  for _ in range(OPTIM_STEPS):
    idx = torch.nonzero(torch.rand(n) < 0.6).squeeze(1)
    m = idx.shape[0]
    steps.append(dict(indexes=idx, visibility=torch.rand(m) * 3 + 0.01, weight=torch.rand(m) * 1.5 + 0.05,
                      basis=torch.randn(m, 2, 2) * 0.3 + 1.5 * torch.eye(2),
                      grads={k: torch.randn_like(v) for k, v in params.items()}))
  return params, groups, steps


def make_optim(out):
  import taichi_splatting.optim.fractional as ref_frac
  import taichi_splatting.optim.visibility_aware as ref_vis
  for ci, cls_name in enumerate(OPTIM_CLASSES):
    params, groups, steps = optim_problem(100 + ci)
    tensors = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
    cls = getattr(ref_vis, cls_name, None) or getattr(ref_frac, cls_name)
    opt = cls([dict(params=[tensors[k]], name=k, **g) for k, g in groups.items()])
    for k, v in params.items():
      out[f"{cls_name}_init_{k}"] = np_(v)
    for k, g in groups.items():
      for extra in ("mask_lr", "point_lr"):
        if extra in g:
          out[f"{cls_name}_{extra}_{k}"] = np_(g[extra])
    for si, st in enumerate(steps):
      for k, t in tensors.items():
        t.grad = st["grads"][k].clone()
        out[f"{cls_name}_s{si}_grad_{k}"] = np_(st["grads"][k])
      out[f"{cls_name}_s{si}_indexes"] = np_(st["indexes"])
      out[f"{cls_name}_s{si}_basis"] = np_(st["basis"])
      if cls_name.startswith("Visibility"):
        out[f"{cls_name}_s{si}_visibility"] = np_(st["visibility"])
        opt.step(indexes=st["indexes"], visibility=st["visibility"], basis=st["basis"])
      elif cls_name.startswith("Sparse"):
        opt.step(indexes=st["indexes"], basis=st["basis"])
      else:
        out[f"{cls_name}_s{si}_weight"] = np_(st["weight"])
        opt.step(indexes=st["indexes"], weight=st["weight"], basis=st["basis"])
      for k, t in tensors.items():
        out[f"{cls_name}_s{si}_param_{k}"] = np_(t)
    for k, t in tensors.items():
      for sk, sv in opt.state[t].items():
        out[f"{cls_name}_state_{k}_{sk}"] = np_(sv)
    print(f"optim {cls_name}: {OPTIM_STEPS} steps, state keys "
          f"{ {k: sorted(opt.state[t].keys()) for k, t in tensors.items()} }")
  out["classes"] = np.array(OPTIM_CLASSES)
  out["num_steps"] = np.array(OPTIM_STEPS)


# ----------------------------------------------------------------------------------------------- ParameterClass + splits
PCLASS_N, PCLASS_KEEP, PCLASS_NEW = 40, 27, 6


def pclass_problem():
  """Tensors, groups and the scripted sequence (steps, a filter, an append) shared by the generator and the tests."""
  torch.manual_seed(77)
  n = PCLASS_N
  tensors = dict(position=torch.randn(n, 2), log_scaling=torch.randn(n, 2) * 0.3, feature=torch.rand(n, 3),
                 z_depth=torch.rand(n, 1))
  groups = dict(position=dict(lr=0.1, type="vector"), log_scaling=dict(lr=0.05, type="scalar"),
                feature=dict(lr=0.02, type="vector"))
  keep = torch.randperm(n)[:PCLASS_KEEP].sort().values
  new = dict(position=torch.randn(PCLASS_NEW, 2), log_scaling=torch.randn(PCLASS_NEW, 2) * 0.3,
             feature=torch.rand(PCLASS_NEW, 3), z_depth=torch.rand(PCLASS_NEW, 1))
  sizes = [n, n, PCLASS_KEEP, PCLASS_KEEP + PCLASS_NEW]   # rows at each of the four optimizer steps
  steps = []
  for m in sizes:
    idx = torch.nonzero(torch.rand(m) < 0.7).squeeze(1)
    steps.append(dict(indexes=idx, grads={k: torch.randn(m, *tensors[k].shape[1:]) for k in groups}))
  return tensors, groups, keep, new, steps


def make_pclass(out):
  """optim/parameter_class.py (ParameterClass: construction, step, boolean / index filtering with optimizer state,
  append with zero state) driven by the reference's SparseAdam (optim/fractional.py + its Taichi step kernels under the
  emulator), and the split operations of misc/renderer2d.py:60-132 with a seeded generator."""
  from tensordict import TensorDict
  from taichi_splatting.optim.parameter_class import ParameterClass
  from taichi_splatting.optim.fractional import SparseAdam
  import taichi_splatting.misc.renderer2d as r2
  tensors, groups, keep, new, steps = pclass_problem()
  for k, v in tensors.items():
    out[f"init_{k}"] = np_(v)
  for k, v in new.items():
    out[f"new_{k}"] = np_(v)
  out["keep"] = np_(keep)
  pc = ParameterClass(TensorDict.from_dict({k: v.clone() for k, v in tensors.items()}, batch_dims=1), groups,
                      optimizer=SparseAdam, betas=(0.9, 0.95), eps=1e-12, bias_correction=True)

  def snapshot(tag, pc):
    for k, v in pc.tensors.items():
      out[f"{tag}_tensor_{k}"] = np_(v)
    for k, st in pc.tensor_state.to_dict().items():
      for sk, sv in st.items():
        out[f"{tag}_state_{k}_{sk}"] = np_(sv)

  def do_step(i, pc):
    st = steps[i]
    out[f"s{i}_indexes"] = np_(st["indexes"])
    for k in groups:
      out[f"s{i}_grad_{k}"] = np_(st["grads"][k])
      pc.tensors[k].grad = st["grads"][k].clone()
    pc.step(indexes=st["indexes"])
    snapshot(f"after_step{i}", pc)

  do_step(0, pc)
  do_step(1, pc)
  pc = pc[keep]                                     # filter: rows AND optimizer state follow
  snapshot("after_filter", pc)
  do_step(2, pc)
  pc = pc.append_tensors(TensorDict.from_dict({k: v.clone() for k, v in new.items()}, batch_dims=1))   # zero state
  snapshot("after_append", pc)
  do_step(3, pc)
  out["state_keys"] = np.array(sorted({k.split("_state_")[1] for k in out if "_state_" in k}))
  print(f"pclass: rows {PCLASS_N} -> filter {PCLASS_KEEP} -> append {PCLASS_KEEP + PCLASS_NEW}; state entries "
        f"{sorted({k.split('_state_')[1] for k in out if k.startswith('after_step3_state_')})}")

  # split operations (misc/renderer2d.py): same generator seed before each call, outputs stored
  torch.manual_seed(5)
  g = random_2d_gaussians(25, (64, 48), num_channels=3, scale_factor=2.0)
  for k in ("position", "z_depth", "log_scaling", "rotation", "alpha_logit", "feature"):
    out[f"split_in_{k}"] = np_(getattr(g, k))
  for name, fn in (("split2", lambda: r2.split_gaussians2d(g, n=2)), ("split3s", lambda: r2.split_gaussians2d(g, n=3, scaling=0.6)),
                   ("uniform2", lambda: r2.uniform_split_gaussians2d(g, n=2)),
                   ("uniform3r", lambda: r2.uniform_split_gaussians2d(g, n=3, random_axis=True, sep=0.5))):
    torch.manual_seed(11)
    res = fn()
    for k in ("position", "z_depth", "log_scaling", "rotation", "alpha_logit", "feature"):
      out[f"{name}_{k}"] = np_(getattr(res, k))
  # (sample_gaussians cannot run in the reference: (N,2,2) @ (N,1,2) is a shape error, renderer2d.py:101-103)
  out["point_basis"] = np_(r2.point_basis(g))
  out["point_covariance"] = np_(r2.point_covariance(g))


MAKERS = {"tile_map": make_tile_map, "pclass": make_pclass, "projection": make_projection, "sh": make_sh, "raster": make_raster,
          "optim": make_optim, "morton": lambda out: make_morton(out)}


# ----------------------------------------------------------------------------------------------- Morton codes
MORTON_CASES = [(0, 300, 0.001, 2.0), (1, 257, 0.01, 50.0), (2, 64, 1e-5, 0.5), (3, 500, 0.05, 1e5)]  # seed, n, resolution, extent


def make_morton(out):
  """code_points64_kernel / code_points32_kernel of misc/morton_sort.py run by the emulator; some points sit far outside
  the grid (clamped to the last cell) and several are duplicated (equal codes: the argsort must be stable)."""
  import taichi_splatting.misc.morton_sort as ms
  for i, (seed, n, resolution, extent) in enumerate(MORTON_CASES):
    torch.manual_seed(seed)
    pts = (torch.rand(n, 3) - 0.3) * extent
    pts[n // 2:n // 2 + 10] = pts[:10]                       # duplicates
    pts = pts.to(torch.float32).contiguous()
    grid = ms.grid_at_resolution(pts, resolution, size=2 ** 20)
    codes64 = torch.empty(n, dtype=torch.uint64)
    ms.code_points64_kernel(grid, pts, codes64)
    grid32 = ms.grid_at_resolution(pts, resolution * 1024, size=2 ** 10)
    codes32 = torch.empty(n, dtype=torch.uint32)
    ms.code_points32_kernel(grid32, pts, codes32)
    c64 = codes64.numpy().copy()
    print(f"morton case {i}: n={n} resolution={resolution} distinct codes {len(np.unique(c64))} "
          f"clamped {(int((c64 == c64.max()).sum()))}")
    out.update({f"c{i}_points": np_(pts), f"c{i}_resolution": np.array(resolution),
                f"c{i}_codes64": c64, f"c{i}_codes32": codes32.numpy().copy(),
                f"c{i}_argsort": np.argsort(c64, kind="stable").astype(np.int32)})   # = cuda_lib.radix_argsort(codes)
  out["num_cases"] = np.array(len(MORTON_CASES))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--only", default="tile_map,projection,sh,raster,optim,morton")
  ap.add_argument("--quick", action="store_true", help="raster: only the two cheapest cases")
  ap.add_argument("--cases", default=None, help="raster2: comma separated case numbers (written to --out)")
  ap.add_argument("--out", default=None, help="output path (default tests/golden/<name>.npz)")
  ap.add_argument("--merge", nargs="*", default=None, help="raster2: join per case .npz files into raster2.npz")
  args = ap.parse_args()
  if args.merge is not None:
    out = {}
    for f in args.merge:
      out.update({k: v for k, v in np.load(f).items()})
    np.savez_compressed(HERE / "raster2.npz", **out)
    print(f"wrote {HERE / 'raster2.npz'} from {len(args.merge)} files, keys {len(out)}")
    return
  for name in args.only.split(","):
    out = {}
    t0 = time.time()
    if name == "raster":
      make_raster(out, quick=args.quick)
    elif name == "raster2":
      make_raster2(out, only=None if args.cases is None else [int(c) for c in args.cases.split(",")])
    else:
      MAKERS[name](out)
    path = Path(args.out) if args.out else HERE / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size / 1024:.0f} KiB, {time.time() - t0:.0f}s)")


if __name__ == "__main__":
  main()
