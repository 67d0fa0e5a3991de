"""Fit 2D gaussians to an image with visibility-weighted sparse optimizers and split / prune densification —
BASELINE.json config 1 as an actual training loop.

Follows taichi_splatting/examples/fit_image_gaussians.py: ``train_epoch`` (:88-150), ``make_epochs`` (:153-169),
``take_n`` / ``randomize_n`` / ``find_split_prune`` / ``split_prune`` (:172-228) and the driver (:232-392) with the same
command line, minus the OpenCV window and Taichi start-up.  Without an image file (or without PIL) it fits a seeded
synthetic target, so that it runs on a box with no data:

  python -m taichi_gaussian_rasterizer_b200.examples.fit_image_gaussians [image] --n 20000 --target 30000 --iters 200

Every step is: project (torch) -> tile map + rasterize forward/backward (CUDA) -> fused optimizer step over the visible
gaussians (CUDA); every epoch the accumulated ``point_heuristic`` (prune cost, split score) drives split / prune on the
ParameterClass, whose optimizer state follows the surviving rows.
"""
import argparse
import json
import math
import time
from pathlib import Path
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from ..data_types import Gaussians2D, RasterConfig
from ..misc.renderer2d import point_basis, project_gaussians2d, uniform_split_gaussians2d
from ..optim import ParameterClass, VisibilityAwareLaProp, VisibilityOptimizer
from ..rasterizer import rasterize
from ..synthetic import random_2d_gaussians


def log_lerp(t, a, b):
  return math.exp(math.log(b) * t + math.log(a) * (1 - t))


def lerp(t, a, b):
  return b * t + a * (1 - t)


def psnr(a, b):
  return 10 * torch.log10(1 / F.mse_loss(a, b))


def check_finite(gaussians: Gaussians2D, what: str):
  for name, t in gaussians.items():
    if not torch.isfinite(t).all():
      raise FloatingPointError(f"{what}: non-finite values in {name}")


def train_epoch(opt, params: ParameterClass, ref_image: torch.Tensor, config: RasterConfig, epoch_size: int = 100,
                opacity_reg: float = 0.0, scale_reg: float = 0.0):
  """``epoch_size`` optimisation steps; returns the last image and the split / prune statistics summed over the epoch
  as ``(prune_cost, split_score)`` (fit_image_gaussians.py:88-150)."""
  h, w = ref_image.shape[:2]
  n, device = params.batch_size[0], params.position.device
  point_heuristic = torch.zeros((n, 2), device=device)
  image = None
  for _ in range(epoch_size):
    opt.zero_grad()
    with torch.enable_grad():
      gaussians = Gaussians2D.from_tensordict(params.tensors)
      raster = rasterize(gaussians2d=project_gaussians2d(gaussians), depth=gaussians.z_depth.clamp(0, 1),
                         features=gaussians.feature, image_size=(w, h), config=config)
      image = raster.image.sigmoid()
      scale = torch.exp(gaussians.log_scaling) / min(w, h)
      loss = (F.mse_loss(image, ref_image) + opacity_reg * gaussians.opacity.mean() + scale_reg * scale.pow(2).mean())
      loss.backward()
    check_finite(gaussians, 'gaussians')

    visibility = raster.visibility
    visible = (visibility > 1e-8).nonzero().squeeze(1)
    basis = point_basis(gaussians[visible]).detach()
    if isinstance(opt, VisibilityOptimizer):
      opt.step(indexes=visible, visibility=visibility[visible], basis=basis)
    else:
      opt.step(indexes=visible, basis=basis)
    with torch.no_grad():  # keep the parameters in range (in place: the optimizer holds these very tensors)
      params.rotation.copy_(F.normalize(params.rotation))
      params.log_scaling.clamp_(min=-5, max=5)
    point_heuristic += raster.point_heuristic
  return image, (point_heuristic[:, 0], point_heuristic[:, 1])


def make_epochs(total_iters: int, first_epoch: int, max_epoch: int):
  """Epoch lengths growing geometrically from ``first_epoch`` to ``max_epoch``; the last one absorbs the remainder."""
  done, epochs = 0, []
  while done < total_iters:
    size = math.ceil(log_lerp(done / total_iters, first_epoch, max_epoch))
    if done + 2 * size > total_iters:
      size = total_iters - done
    done += size
    epochs.append(size)
  return epochs


def take_n(t: torch.Tensor, n: int, descending: bool = False) -> torch.Tensor:
  """Mask of the ``n`` largest (``descending``) or smallest values of ``t``."""
  mask = torch.zeros_like(t, dtype=torch.bool)
  mask[torch.argsort(t, descending=descending)[:max(n, 0)]] = True
  return mask


def randomize_n(t: torch.Tensor, n: int) -> torch.Tensor:
  """Mask of ``n`` entries drawn without replacement with probability proportional to ``t``."""
  mask = torch.zeros_like(t, dtype=torch.bool)
  if n > 0:
    mask[torch.multinomial(F.normalize(t, dim=0), n, replacement=False)] = True
  return mask


def find_split_prune(n: int, target: int, n_prune: int, prune_cost: torch.Tensor, densify_score: torch.Tensor):
  """Prune the ``n_prune`` cheapest points and split as many of the best-scoring ones as it takes to reach ``target``
  (each split adds one point); a point selected for both is left alone (fit_image_gaussians.py:190-202)."""
  prune_mask = take_n(prune_cost, n_prune, descending=False)
  pruned = int(prune_mask.sum().item())
  split_mask = take_n(densify_score, max(0, (target - n) + pruned), descending=True)
  both = split_mask & prune_mask
  return split_mask ^ both, prune_mask ^ both


def split_prune(params: ParameterClass, t: float, target: int, prune_rate: float,
                split_heuristic: Tuple[torch.Tensor, torch.Tensor]):
  """One densification round on the ParameterClass; children start with zero optimizer state
  (fit_image_gaussians.py:204-228)."""
  n = params.batch_size[0]
  prune_cost, split_score = split_heuristic
  split_mask, prune_mask = find_split_prune(n=n, target=target, n_prune=int(prune_rate * n * (1 - t)),
                                            prune_cost=prune_cost, densify_score=split_score)
  to_split = params[split_mask]
  children = uniform_split_gaussians2d(Gaussians2D.from_tensordict(to_split.tensors.detach()), random_axis=True)
  child_state = to_split.tensor_state.new_zeros(to_split.batch_size[0], 2)
  params = params[~(split_mask | prune_mask)]
  params = params.append_tensors(children.to_tensordict(), child_state.reshape(children.batch_size))
  return params, dict(split=int(split_mask.sum().item()), prune=int(prune_mask.sum().item()))


def synthetic_target(image_size: Tuple[int, int], seed: int, device) -> torch.Tensor:
  """A smooth seeded colour image (sum of a few hundred soft blobs) standing in for a photograph."""
  w, h = image_size
  g = torch.Generator().manual_seed(seed)
  ys, xs = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing='ij')
  img = torch.zeros(h, w, 3)
  for _ in range(64):
    cx, cy, r = torch.rand(3, generator=g).tolist()
    colour = torch.rand(3, generator=g)
    blob = torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * (0.02 + 0.15 * r) ** 2))
    img = img * (1 - 0.7 * blob[..., None]) + 0.7 * blob[..., None] * colour
  return img.to(device)


def load_image(path: Optional[str], image_size: Tuple[int, int], seed: int, device) -> torch.Tensor:
  if path is None:
    return synthetic_target(image_size, seed, device)
  from PIL import Image
  import numpy as np
  rgb = np.asarray(Image.open(path).convert('RGB'), dtype=np.float32) / 255
  return torch.from_numpy(rgb).to(device)


def default_parameter_groups(max_lr: float) -> Dict[str, Dict]:
  return dict(position=dict(lr=max_lr, type='local_vector'), log_scaling=dict(lr=0.1), rotation=dict(lr=1.0),
              alpha_logit=dict(lr=0.1), feature=dict(lr=0.1, type='vector'))


def fit(ref_image: torch.Tensor, n: int, target: Optional[int], iters: int, config: RasterConfig, seed: int = 0,
        max_lr: float = 0.5, min_lr: float = 0.1, epoch: int = 8, max_epoch: int = 32, prune_rate: float = 0.025,
        opacity_reg: float = 1e-5, scale_reg: float = 0.1, log=None):
  """The reference driver's loop (fit_image_gaussians.py:262-392) as a function; returns (params, history)."""
  device = ref_image.device
  h, w = ref_image.shape[:2]
  torch.manual_seed(seed)
  gaussians = random_2d_gaussians(n, (w, h), alpha_range=(0.5, 1.0), scale_factor=0.5).to(device)
  params = ParameterClass(gaussians.to_tensordict(), default_parameter_groups(max_lr), optimizer=VisibilityAwareLaProp,
                          vis_smooth=0.1, vis_beta=0.8, betas=(0.9, 0.9), eps=1e-16, bias_correction=True)
  history, iteration = [], 0
  with torch.no_grad():
    for epoch_size in make_epochs(iters, epoch, max_epoch):
      t = (iteration + epoch_size * 0.5) / iters
      params.set_learning_rate(position=log_lerp(t, max_lr, min_lr))
      start = time.time()
      image, heuristic = train_epoch(params.optimizer, params, ref_image, config=config, epoch_size=epoch_size,
                                     opacity_reg=opacity_reg, scale_reg=scale_reg)
      torch.cuda.synchronize(device)
      seconds = time.time() - start
      entry = dict(iteration=iteration + epoch_size, n=params.batch_size[0], psnr=float(psnr(ref_image, image)),
                   iters_per_s=epoch_size / seconds)
      if target is not None:
        params, counts = split_prune(params, t, target, prune_rate, heuristic)
        entry.update(counts)
      iteration += epoch_size
      history.append(entry)
      if log:
        log(entry)
  return params, history


def parse_args(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('image_file', type=str, nargs='?', default=None)
  ap.add_argument('--seed', type=int, default=0)
  ap.add_argument('--tile_size', type=int, default=16)
  ap.add_argument('--pixel_tile', type=str, help='pixel tile of the backward pass, default "2,2"')
  ap.add_argument('--n', type=int, default=1000)
  ap.add_argument('--target', type=int, default=None)
  ap.add_argument('--prune', action='store_true', help='enable pruning (equivalent to --target=n)')
  ap.add_argument('--iters', type=int, default=2000)
  ap.add_argument('--max_lr', type=float, default=0.5)
  ap.add_argument('--min_lr', type=float, default=0.1)
  ap.add_argument('--epoch', type=int, default=8, help='base epoch size (increases with t)')
  ap.add_argument('--max_epoch', type=int, default=32)
  ap.add_argument('--prune_rate', type=float, default=0.025)
  ap.add_argument('--opacity_reg', type=float, default=0.00001)
  ap.add_argument('--scale_reg', type=float, default=0.1)
  ap.add_argument('--antialias', action='store_true')
  ap.add_argument('--image_size', type=int, nargs=2, default=(1024, 1024), help='size of the synthetic target (w h)')
  ap.add_argument('--write_image', type=Path, default=None)
  args = ap.parse_args(argv)
  if args.pixel_tile:
    args.pixel_tile = tuple(map(int, args.pixel_tile.split(',')))
  if args.prune and args.target is None:
    args.target = args.n
  return args


def main(argv=None):
  args = parse_args(argv)
  device = torch.device('cuda:0')
  ref_image = load_image(args.image_file, tuple(args.image_size), args.seed, device)
  config = RasterConfig(compute_point_heuristic=True, compute_visibility=True, tile_size=args.tile_size,
                        blur_cov=0.0 if args.antialias else 0.3, antialias=args.antialias,
                        pixel_stride=args.pixel_tile or (2, 2))
  params, history = fit(ref_image, n=args.n, target=args.target, iters=args.iters, config=config, seed=args.seed,
                        max_lr=args.max_lr, min_lr=args.min_lr, epoch=args.epoch, max_epoch=args.max_epoch,
                        prune_rate=args.prune_rate, opacity_reg=args.opacity_reg, scale_reg=args.scale_reg,
                        log=lambda e: print(json.dumps(e)))
  if args.write_image is not None:
    from PIL import Image
    g = Gaussians2D.from_tensordict(params.tensors.detach())
    out = rasterize(project_gaussians2d(g), g.z_depth.clamp(0, 1), g.feature,
                    image_size=(ref_image.shape[1], ref_image.shape[0]), config=config).image.sigmoid()
    Image.fromarray((out.clamp(0, 1) * 255).byte().cpu().numpy()).save(args.write_image)
  return history


if __name__ == '__main__':
  main()
