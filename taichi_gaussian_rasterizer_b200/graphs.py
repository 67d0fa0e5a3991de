"""A training (or rendering) step replayed from ONE CUDA graph (extension; the reference has no counterpart — its
render path synchronises the device four times per view, so it cannot be captured at all).

What makes a step capturable here is ``render_gaussians(..., overlap_capacity=)`` (renderer.py): the visible count and
the overlap total stay on the device, so forward + loss + backward of any number of views — issued on several CUDA
streams if the caller likes (distributed.run_views), gradients accumulating in a ``GradientBucket`` — is a fixed
sequence of kernels whose sizes do not depend on the data.  ``CapturedStep`` wraps the warm-up / capture / replay
protocol of ``torch.cuda.graph``; ``overlap_capacity_for`` sizes the capacity from the views at hand.
"""
from typing import Callable, Iterable, Optional

import torch

from .data_types import Gaussians3D, RasterConfig
from .perspective import CameraParams


def overlap_capacity_for(gaussians: Gaussians3D, cameras: Iterable[CameraParams], config: RasterConfig,
                         slack: float = 1.25, use_depth16: bool = False) -> int:
  """An overlap capacity for ``render_gaussians(..., overlap_capacity=)`` that covers every camera of ``cameras`` with
  ``slack`` to spare: the largest tile-overlap count K among them (each measured with one projection + tile mapping,
  which reads K back — call this outside the step), times ``slack``.  A scene that moves may outgrow it: pass
  ``overlap_total_out`` and compare now and then."""
  from .mapper.tile_mapper import map_to_tiles
  from .perspective.projection import project_to_image
  from .torch_lib.projection import ndc_depth
  worst = 0
  with torch.no_grad():
    for cam in cameras:
      g2d, depths, _ = project_to_image(gaussians, cam, config)
      o2p, _ = map_to_tiles(g2d, ndc_depth(depths, cam.near_plane, cam.far_plane), cam.image_size, config,
                            use_depth16=use_depth16)
      worst = max(worst, int(o2p.shape[0]))
  return int(worst * slack) + 4096


class CapturedStep:
  """``fn()`` captured once, replayed many times.

  ``fn`` must not synchronise the host (no ``.item()``, no default ``render_gaussians``: pass ``overlap_capacity``),
  must read its inputs from tensors that keep their address (update them in place between replays), and leaves its
  results — the gradients (``.grad`` tensors or a GradientBucket), whatever it returns — in tensors that every replay
  rewrites.  It is run ``warmup`` times eagerly on a side stream first (allocator pools, lazy initialisation), then
  captured; ``fn`` can ask ``CapturedStep.capturing()`` whether it is being captured (e.g. to skip waits on events of
  a previous eager step).  Gradients that are ``None`` at capture time are ASSIGNED by every replay (no accumulation
  pass); gradients that exist are accumulated into, so zero them inside ``fn``."""

  def __init__(self, fn: Callable[[], object], device: Optional[torch.device] = None, warmup: int = 2,
               before_capture: Optional[Callable[[], None]] = None):
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      for _ in range(warmup):
        if before_capture is not None:
          before_capture()
        fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    if before_capture is not None:
      before_capture()
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph):
      self.result = fn()
    self._keep_alive = fn   # what the graph reads must outlive it: the replay dereferences the captured addresses

  @staticmethod
  def capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()

  def replay(self):
    """Enqueue the whole step (one cudaGraphLaunch) on the current stream; returns what ``fn`` returned at capture
    time (tensors of the graph's memory pool, rewritten by this replay)."""
    self.graph.replay()
    return self.result
