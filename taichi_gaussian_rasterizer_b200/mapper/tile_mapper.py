"""Maps 2D gaussians to the image tiles they overlap, sorted front to back per tile.

Operator surface of taichi_splatting/mapper/tile_mapper.py: ``map_to_tiles`` (:202-223) and
``pad_to_tile`` (:18-22).  The reference counts overlaps per gaussian, scans, emits one 64 bit
(tile, depth) key per overlap and radix-sorts the K keys on 48 bits (6 passes over 12 B pairs).  The
same total order (tile, depth bits, gaussian index) is produced here "depth first" with a third of the
traffic (csrc/geom_kernels.cu, csrc/scan_sort.cu):

  1. stable radix sort of the V gaussians by their 32 bit depth key (f32 bits, or depth16) -> ``perm``;
  2. OBB overlap count in that order -> exclusive scan -> emit bare tile ids + gaussian index in that
     order (each tile receives its gaussians already sorted by (depth, index));
  3. stable radix sort of the K (tile id, index) pairs on the ceil(log2 T) tile bits only (2 passes);
  4. tile range detection.

``map_to_tiles_staged`` keeps the reference's stage order (count -> scan -> 64 bit keys -> sort ->
ranges) on the same kernels; both are bit-identical (tests/test_gpu_tile_mapper.py).  The only host
read-back is the overlap total K, which sizes the key buffers (the reference synchronises the whole
device twice here, cuda_lib/full_cumsum.cu:45 and cuda_lib/radix_sort_pairs.cu:27).
"""
import ctypes
import math
from numbers import Integral

import torch
from beartype import beartype
from beartype.typing import Optional, Tuple

from .. import _native as N
from ..cuda_lib import full_cumsum_device, radix_sort_pairs, radix_sort_pairs_counted
from ..data_types import RasterConfig

MAX_TILE = 65535  # 16 bit tile id inside the sorted key bits (tile_mapper.py:29)


def _pinned_total(device) -> torch.Tensor:
  """A pinned int32 word for the asynchronous read-back of the overlap total: the next slot of a per-device ring
  (_native.PinnedWords), so overlapping read-backs on one stream never share a word."""
  return N.pinned_words.take(device)


def pad_to_tile(image_size: Tuple[Integral, Integral], tile_size: int):
  def pad(x):
    return int(math.ceil(x / tile_size) * tile_size)
  return tuple(pad(x) for x in image_size)


def tile_shape(image_size, tile_size):
  padded = pad_to_tile(image_size, tile_size)
  return (padded[1] // tile_size, padded[0] // tile_size)


def _tile_params(n, image_size, config, use_depth16):
  return N.GsTileParams(int(image_size[0]), int(image_size[1]), config.tile_size, int(use_depth16), n,
                        float(config.alpha_threshold))


def _check_inputs(gaussians, depth, image_size, config):
  assert gaussians.ndim == 2 and gaussians.shape[1] == 7, f"gaussians must be Nx7 got {gaussians.shape}"
  assert depth.ndim == 2 and depth.shape[1] == 1, f"depths must be Nx1, got {depth.shape}"
  assert gaussians.shape[0] == depth.shape[0]
  N.require_cuda(gaussians, depth)
  shape = tile_shape(image_size, config.tile_size)
  assert shape[0] * shape[1] < MAX_TILE, \
    f"tile dimensions {shape} for image size {image_size} exceed maximum tile count (16 bit id), try increasing tile_size"
  return shape


@beartype
def map_to_tiles(gaussians: torch.Tensor, depth: torch.Tensor,
                 image_size: Tuple[Integral, Integral],
                 config: RasterConfig,
                 use_depth16: bool = False,
                 overlap_capacity: Optional[int] = None,
                 overlap_total_out: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
  """ maps gaussians to tiles, sorted by depth (front to back):
    Parameters:
     gaussians: (N, 7) torch.Tensor of packed gaussians (float32)
     depth: (N, 1)  torch.Tensor of depths (float32, >= 0)
     image_size: (2, ) tuple of ints, (width, height)
     config: RasterConfig (tile_size, alpha_threshold)

    Returns:
     overlap_to_point: (K, ) int32, maps overlap index to point index
     tile_ranges: (TH, TW, 2) int32, maps tile index to its [start, end) range of overlap indices

    ``overlap_capacity`` (extension): bound on the number of overlaps K.  ``overlap_to_point`` then has exactly that
    many rows (those past K are undefined), K never travels to the host — no synchronisation at all, so a training step
    can be captured in a CUDA graph — and overlaps beyond the capacity are DROPPED: pass a 1-element int32 CUDA tensor
    as ``overlap_total_out`` to receive K and compare it with the capacity now and then.
  """
  if overlap_capacity is not None:
    return _map_to_tiles_capped(gaussians, depth, image_size, config, use_depth16, int(overlap_capacity),
                                overlap_total_out)
  return _map_to_tiles(gaussians, depth, image_size, config, use_depth16, ndc_range=None)


def _map_to_tiles_capped(gaussians, depth, image_size, config, use_depth16, capacity, total_out=None):
  """The depth-first mapping with capacity-sized key buffers and the overlap total left on the device."""
  shape = _check_inputs(gaussians, depth, image_size, config)
  assert capacity >= 0
  with torch.no_grad():
    device = gaussians.device
    g = gaussians.detach().to(torch.float32).contiguous()
    d = depth.detach().to(torch.float32).contiguous()
    n = g.shape[0]
    stream = N.stream_ptr(device)
    p = _tile_params(n, image_size, config, use_depth16)
    tile_ranges = torch.empty((*shape, 2), dtype=torch.int32, device=device)
    overlap_to_point = torch.empty((capacity,), dtype=torch.int32, device=device)
    if n == 0 or capacity == 0:
      tile_ranges.zero_()
      if total_out is not None:
        total_out.zero_()
      return overlap_to_point, tile_ranges
    depth_keys = torch.empty((n,), dtype=torch.int32, device=device)
    iota = torch.empty((n,), dtype=torch.int32, device=device)
    N.call("gs_depth_keys", ctypes.byref(p), N.ptr(d), ctypes.c_double(0.0), ctypes.c_double(0.0), N.ptr(depth_keys),
           N.ptr(iota), stream)
    _, perm = radix_sort_pairs(depth_keys, iota, 0, 16 if use_depth16 else 32)
    counts = torch.empty((n,), dtype=torch.int32, device=device)
    masks = torch.empty((n,), dtype=torch.int64, device=device)
    N.call("gs_tile_count_perm", ctypes.byref(p), N.ptr(g), N.ptr(perm), N.ptr(counts), N.ptr(masks), stream)
    cum = full_cumsum_device(counts)
    total_dev = cum[n:]                      # K, on the device
    if total_out is not None:
      total_out.copy_(total_dev.reshape(total_out.shape))
    tile_ids = torch.empty((capacity,), dtype=torch.int32, device=device)
    values = torch.empty((capacity,), dtype=torch.int32, device=device)
    N.call("gs_tile_emit_tiles_capped", ctypes.byref(p), N.ptr(g), N.ptr(perm), N.ptr(cum), N.ptr(masks),
           ctypes.c_int64(capacity), N.ptr(tile_ids), N.ptr(values), stream)
    tile_bits = max(1, (shape[0] * shape[1] - 1).bit_length())
    tile_ids, overlap_to_point = radix_sort_pairs_counted(tile_ids, values, total_dev, 0, tile_bits)
    N.call("gs_find_ranges_tiles_counted", ctypes.byref(p), ctypes.c_int64(capacity), N.ptr(total_dev), N.ptr(tile_ids),
           N.ptr(tile_ranges), stream)
    return overlap_to_point, tile_ranges


def launch_depth_order_counted(depth_capacity, count_device, image_size, config, use_depth16=False, ndc_range=None):
  """Stage 1 of the depth-first mapping (depth keys + their stable sort) for a visible set whose SIZE is still on the
  device: ``depth_capacity`` (N, 1) f32 holds the depths in its first ``count_device[0]`` rows.  Returns the (N,)
  int32 permutation buffer, valid in its first count rows; hand ``perm[:V]`` to ``_map_to_tiles(depth_order=...)``.
  render_gaussians enqueues this before the host reads the visible count, so that the GPU has work across that
  read-back and the host's launch latency after it."""
  cap = depth_capacity.shape[0]
  near, far = (float(ndc_range[0]), float(ndc_range[1])) if ndc_range is not None else (0.0, 0.0)
  device = depth_capacity.device
  p = _tile_params(cap, image_size, config, use_depth16)
  depth_keys = torch.empty((cap,), dtype=torch.int32, device=device)
  iota = torch.empty((cap,), dtype=torch.int32, device=device)
  if cap == 0:
    return iota
  N.call("gs_depth_keys_counted", ctypes.byref(p), N.ptr(depth_capacity), ctypes.c_double(near), ctypes.c_double(far),
         N.ptr(count_device), N.ptr(depth_keys), N.ptr(iota), N.stream_ptr(device))
  _, perm = radix_sort_pairs_counted(depth_keys, iota, count_device, 0, 16 if use_depth16 else 32)
  return perm


class MapperFront:
  """What launch_mapper_front_counted leaves on the device (capacity-sized buffers, valid in their first V rows) and
  the pending read-back of the overlap total."""
  __slots__ = ("perm", "counts", "masks", "cum", "host_total", "total_ready", "total_dev")


def launch_mapper_front_counted(points_capacity, depth_capacity, count_device, image_size, config, use_depth16=False,
                                ndc_range=None, read_back=True) -> MapperFront:
  """Everything of the depth-first mapping that does not need the overlap total, for a visible set whose SIZE is still
  on the device: depth keys + their sort, the overlap count in depth order, its scan, and the asynchronous copy of the
  total K into a pinned word.  render_gaussians enqueues this right behind the projection kernel: about 0.3 ms of GPU
  work (3 M gaussians) that covers the host's read-back of the visible count AND the Python between that and the
  first launch that needs K; ``_map_to_tiles(front=...)`` then only waits for K, emits and sorts the tile ids."""
  f = MapperFront()
  cap = depth_capacity.shape[0]
  device = depth_capacity.device
  f.perm = launch_depth_order_counted(depth_capacity, count_device, image_size, config, use_depth16, ndc_range)
  p = _tile_params(cap, image_size, config, use_depth16)
  stream = N.stream_ptr(device)
  f.counts = torch.empty((cap,), dtype=torch.int32, device=device)
  f.masks = torch.empty((cap,), dtype=torch.int64, device=device)
  f.cum = torch.empty((cap + 1,), dtype=torch.int32, device=device)
  total_dev = torch.empty((1,), dtype=torch.int32, device=device)
  lib = N.lib()
  if cap > 0:
    N.call("gs_tile_count_perm_counted", ctypes.byref(p), N.ptr(points_capacity), N.ptr(f.perm), N.ptr(count_device),
           N.ptr(f.counts), N.ptr(f.masks), stream)
  ws = N.workspace(lib.gs_full_cumsum_workspace_bytes(cap, 4), device)
  N.call("gs_full_cumsum_counted", ctypes.c_int64(cap), ctypes.c_int32(4), N.ptr(count_device), N.ptr(f.counts),
         N.ptr(f.cum), N.ptr(total_dev), N.ptr(ws), ctypes.c_size_t(ws.numel()), stream)
  f.total_dev = total_dev
  f.host_total = f.total_ready = None
  if read_back:   # not for map_front_to_tiles_capped: the total never leaves the device there
    f.host_total = _pinned_total(device)
    f.host_total.copy_(total_dev, non_blocking=True)
    f.total_ready = torch.cuda.Event()
    f.total_ready.record(torch.cuda.current_stream(device))
  return f


def map_front_to_tiles_capped(gaussians_capacity, count_device, front: MapperFront, image_size, config, capacity,
                              total_out=None):
  """The back half of the depth-first mapping for a visible set whose size AND overlap total both stay on the device
  (render_gaussians(..., overlap_capacity=)): ``gaussians_capacity`` (N, 7) holds the packed gaussians in its first
  ``count_device[0]`` rows, ``front`` is their MapperFront (launch_mapper_front_counted(..., read_back=False)).
  Returns (overlap_to_point (capacity,), tile_ranges (TH, TW, 2)); overlaps beyond the capacity are dropped, K goes to
  ``total_out`` (1-element int32 CUDA tensor) when given."""
  capacity = int(capacity)
  assert capacity >= 0
  shape = tile_shape(image_size, config.tile_size)
  assert shape[0] * shape[1] < MAX_TILE, \
    f"tile dimensions {shape} for image size {image_size} exceed maximum tile count (16 bit id), try increasing tile_size"
  with torch.no_grad():
    device = gaussians_capacity.device
    g = gaussians_capacity.detach()
    assert g.dtype == torch.float32 and g.is_contiguous()
    n = g.shape[0]
    stream = N.stream_ptr(device)
    p = _tile_params(n, image_size, config, False)
    tile_ranges = torch.empty((*shape, 2), dtype=torch.int32, device=device)
    overlap_to_point = torch.empty((capacity,), dtype=torch.int32, device=device)
    if total_out is not None:
      total_out.copy_(front.total_dev.reshape(total_out.shape))
    if n == 0 or capacity == 0:
      tile_ranges.zero_()
      return overlap_to_point, tile_ranges
    tile_ids = torch.empty((capacity,), dtype=torch.int32, device=device)
    values = torch.empty((capacity,), dtype=torch.int32, device=device)
    N.call("gs_tile_emit_tiles_capped_counted", ctypes.byref(p), N.ptr(g), N.ptr(front.perm), N.ptr(front.cum),
           N.ptr(front.masks), N.ptr(count_device), ctypes.c_int64(capacity), N.ptr(tile_ids), N.ptr(values), stream)
    tile_bits = max(1, (shape[0] * shape[1] - 1).bit_length())
    tile_ids, overlap_to_point = radix_sort_pairs_counted(tile_ids, values, front.total_dev, 0, tile_bits)
    N.call("gs_find_ranges_tiles_counted", ctypes.byref(p), ctypes.c_int64(capacity), N.ptr(front.total_dev),
           N.ptr(tile_ids), N.ptr(tile_ranges), stream)
    return overlap_to_point, tile_ranges


def _map_to_tiles(gaussians, depth, image_size, config, use_depth16=False, ndc_range=None, depth_order=None,
                  before_total_sync=None, front=None):
  """map_to_tiles; with ``ndc_range=(near, far)`` the ``depth`` column is LINEAR camera depth and the sort depth is
  its NDC value, formed inside the key kernel with torch's own f32 operation sequence (bit-identical keys to
  ``map_to_tiles(g, ndc_depth(depth, near, far), ...)`` — tests/test_gpu_tile_mapper.py), which saves the
  render path four elementwise launches per frame (render_projected).  ``depth_order``: the (n,) permutation already
  produced by launch_depth_order_counted for exactly these depths.  ``before_total_sync()``: called once, after the
  overlap scan is enqueued and before the host waits for the overlap total — the caller's chance to enqueue work that
  does not depend on the tile map, so that the GPU is busy across that read-back.  ``front``: the MapperFront of
  launch_mapper_front_counted for exactly these gaussians (then ``depth_order`` is not needed)."""
  shape = _check_inputs(gaussians, depth, image_size, config)
  near, far = (float(ndc_range[0]), float(ndc_range[1])) if ndc_range is not None else (0.0, 0.0)
  with torch.no_grad():
    device = gaussians.device
    g = gaussians.detach().to(torch.float32).contiguous()   # the tile mapper is f32 (tile_mapper.py:12)
    d = depth.detach().to(torch.float32).contiguous()
    n = g.shape[0]
    stream = N.stream_ptr(device)
    p = _tile_params(n, image_size, config, use_depth16)
    tile_ranges = torch.empty((*shape, 2), dtype=torch.int32, device=device)

    total = 0
    if n > 0 and front is not None:
      perm, masks, cum = front.perm, front.masks, front.cum
      tile_bits = max(1, (shape[0] * shape[1] - 1).bit_length())
      if before_total_sync is not None:
        before_total_sync()
        before_total_sync = None
      front.total_ready.synchronize()
      total = int(front.host_total.item())
    elif n > 0:
      if depth_order is not None:
        assert depth_order.shape == (n,) and depth_order.dtype == torch.int32 and depth_order.is_contiguous()
        perm = depth_order
      else:
        depth_keys = torch.empty((n,), dtype=torch.int32, device=device)   # bit patterns of u32 keys
        iota = torch.empty((n,), dtype=torch.int32, device=device)
        N.call("gs_depth_keys", ctypes.byref(p), N.ptr(d), ctypes.c_double(near), ctypes.c_double(far),
               N.ptr(depth_keys), N.ptr(iota), stream)
        _, perm = radix_sort_pairs(depth_keys, iota, 0, 16 if use_depth16 else 32)
      counts = torch.empty((n,), dtype=torch.int32, device=device)
      masks = torch.empty((n,), dtype=torch.int64, device=device)   # per slot tile bit masks: count pass -> emit pass
      N.call("gs_tile_count_perm", ctypes.byref(p), N.ptr(g), N.ptr(perm), N.ptr(counts), N.ptr(masks), stream)
      cum = full_cumsum_device(counts)
      # host read-back of K: asynchronous copy into a pinned word + event wait (no stream-wide synchronisation, no
      # pageable staging as `.item()` on a device tensor does)
      host_total = _pinned_total(device)
      host_total.copy_(cum[n:], non_blocking=True)
      ready = torch.cuda.Event()
      ready.record(torch.cuda.current_stream(device))
      tile_bits = max(1, (shape[0] * shape[1] - 1).bit_length())
      if before_total_sync is not None:
        before_total_sync()
        before_total_sync = None
      ready.synchronize()
      total = int(host_total.item())

    if total > 0:
      tile_ids = torch.empty((total,), dtype=torch.int32, device=device)
      values = torch.empty((total,), dtype=torch.int32, device=device)
      N.call("gs_tile_emit_tiles", ctypes.byref(p), N.ptr(g), N.ptr(perm), N.ptr(cum), N.ptr(masks), N.ptr(tile_ids),
             N.ptr(values), stream)
      tile_ids, overlap_to_point = radix_sort_pairs(tile_ids, values, 0, tile_bits)
    else:
      tile_ids = None
      overlap_to_point = torch.empty((0,), dtype=torch.int32, device=device)

    N.call("gs_find_ranges_tiles", ctypes.byref(p), ctypes.c_int64(total), N.ptr(tile_ids), N.ptr(tile_ranges), stream)
    if before_total_sync is not None:   # nothing in view: the callback still runs exactly once
      before_total_sync()
    return overlap_to_point, tile_ranges


@beartype
def map_to_tiles_staged(gaussians: torch.Tensor, depth: torch.Tensor,
                        image_size: Tuple[Integral, Integral],
                        config: RasterConfig,
                        use_depth16: bool = False
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
  """ map_to_tiles in the reference's stage order (one 64 bit key per overlap, one K-sized sort); same result.
    Parameters:
     gaussians: (N, 7) torch.Tensor of packed gaussians (float32)
     depth: (N, 1)  torch.Tensor of depths (float32, >= 0)
     image_size: (2, ) tuple of ints, (width, height)
     config: RasterConfig (tile_size, alpha_threshold)

    Returns:
     overlap_to_point: (K, ) int32, maps overlap index to point index
     tile_ranges: (TH, TW, 2) int32, maps tile index to its [start, end) range of overlap indices
  """
  assert gaussians.ndim == 2 and gaussians.shape[1] == 7, f"gaussians must be Nx7 got {gaussians.shape}"
  assert depth.ndim == 2 and depth.shape[1] == 1, f"depths must be Nx1, got {depth.shape}"
  assert gaussians.shape[0] == depth.shape[0]
  N.require_cuda(gaussians, depth)

  shape = tile_shape(image_size, config.tile_size)
  assert shape[0] * shape[1] < MAX_TILE, \
    f"tile dimensions {shape} for image size {image_size} exceed maximum tile count (16 bit id), try increasing tile_size"

  with torch.no_grad():
    device = gaussians.device
    g = gaussians.detach().to(torch.float32).contiguous()   # the tile mapper is f32 (tile_mapper.py:12)
    d = depth.detach().to(torch.float32).contiguous()
    n = g.shape[0]
    lib = N.lib()
    stream = N.stream_ptr(device)
    p = _tile_params(n, image_size, config, use_depth16)
    tile_ranges = torch.empty((*shape, 2), dtype=torch.int32, device=device)

    total = 0
    if n > 0:
      counts = torch.empty((n,), dtype=torch.int32, device=device)
      N.call("gs_tile_count", ctypes.byref(p), N.ptr(g), N.ptr(counts), stream)
      cum = full_cumsum_device(counts)
      # host read-back of K: asynchronous copy into a pinned word + event wait (no stream-wide synchronisation, no
      # pageable staging as `.item()` on a device tensor does)
      host_total = _pinned_total(device)
      host_total.copy_(cum[n:], non_blocking=True)
      ready = torch.cuda.Event()
      ready.record(torch.cuda.current_stream(device))
      ready.synchronize()
      total = int(host_total.item())

    if total > 0:
      key_dtype = torch.int32 if use_depth16 else torch.int64   # bit patterns of u32 / u64 keys
      keys = torch.empty((total,), dtype=key_dtype, device=device)
      values = torch.empty((total,), dtype=torch.int32, device=device)
      N.call("gs_tile_emit_keys", ctypes.byref(p), N.ptr(g), N.ptr(d), N.ptr(cum), N.ptr(keys), N.ptr(values),
                                    stream)
      # depth bits + only as many tile-id bits as there are tiles (same order as sorting all 16)
      tile_bits = max(1, (shape[0] * shape[1] - 1).bit_length())
      end_bit = (16 if use_depth16 else 32) + tile_bits
      keys, overlap_to_point = radix_sort_pairs(keys, values, 0, end_bit)
    else:
      keys = None
      overlap_to_point = torch.empty((0,), dtype=torch.int32, device=device)

    N.call("gs_find_ranges", ctypes.byref(p), ctypes.c_int64(total), N.ptr(keys), N.ptr(tile_ranges), stream)
    return overlap_to_point, tile_ranges
