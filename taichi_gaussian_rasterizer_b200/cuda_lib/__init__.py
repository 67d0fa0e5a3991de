"""Scan and sort primitives with the surface of taichi_splatting/cuda_lib/__init__.py:16-43.

``full_cumsum`` (cuda_lib/full_cumsum.cu:16-67: CUB ExclusiveSum + device sync) is a single-pass
chained scan here, ``radix_sort_pairs`` (cuda_lib/radix_sort_pairs.cu:9-69: CUB SortPairs +
device sync) a hand-written onesweep LSD radix sort (csrc/scan_sort.cu).  Neither synchronises
the device; ``full_cumsum`` reads the total back because its signature returns a Python int.
"""
import ctypes

import torch
from beartype.typing import Tuple

from .. import _native as N

_KEY_BYTES = {torch.int32: 4, torch.uint32: 4, torch.int64: 8, torch.uint64: 8}


def check_cuda(name, arg):
  N.require_cuda(arg)


def full_cumsum_device(x: torch.Tensor) -> torch.Tensor:
  """Exclusive scan with the total appended: out (n+1,), out[n] = sum(x).  No host sync."""
  check_cuda("full_cumsum", x)
  assert x.dtype in (torch.int32, torch.int64) and x.dim() == 1, "full_cumsum: 1-D int32 / int64 only"
  n = x.shape[0]
  out = torch.empty((n + 1,), dtype=x.dtype, device=x.device)
  lib = N.lib()
  eb = x.element_size()
  ws = N.workspace(lib.gs_full_cumsum_workspace_bytes(n, eb), x.device)
  N.call("gs_full_cumsum", ctypes.c_int64(n), ctypes.c_int32(eb), N.ptr(x.contiguous()), N.ptr(out), N.ptr(ws),
                             ctypes.c_size_t(ws.numel()), N.stream_ptr(x.device))
  return out


def full_cumsum(x: torch.Tensor) -> Tuple[torch.Tensor, int]:
  check_cuda("full_cumsum", x)
  if x.shape[0] == 0:
    return x.new_zeros((1,)), 0
  out = full_cumsum_device(x)
  return out, int(out[-1].item())


def radix_sort_pairs(keys: torch.Tensor, values: torch.Tensor, start_bit=0, end_bit=None):
  """Stable sort of (key, int32 value) pairs on key bits [start_bit, end_bit)."""
  check_cuda("keys", keys)
  check_cuda("values", values)
  assert keys.dtype in _KEY_BYTES, f"keys must be a 32 or 64 bit integer tensor, got {keys.dtype}"
  assert values.dtype == torch.int32, f"values must be int32, got {values.dtype}"
  assert keys.shape == values.shape and keys.dim() == 1
  kb = _KEY_BYTES[keys.dtype]
  if end_bit is None or end_bit < 0:
    end_bit = kb * 8
  n = keys.shape[0]
  keys_out, values_out = torch.empty_like(keys), torch.empty_like(values)
  if n == 0:
    return keys_out, values_out
  lib = N.lib()
  ws = N.workspace(lib.gs_radix_sort_pairs_workspace_bytes(n, kb, start_bit, end_bit), keys.device)
  N.call("gs_radix_sort_pairs", ctypes.c_int64(n), ctypes.c_int32(kb), N.ptr(keys.contiguous()),
                                  N.ptr(values.contiguous()), N.ptr(keys_out), N.ptr(values_out),
                                  ctypes.c_int32(start_bit), ctypes.c_int32(end_bit), N.ptr(ws),
                                  ctypes.c_size_t(ws.numel()), N.stream_ptr(keys.device))
  return keys_out, values_out


def radix_sort_pairs_counted(keys: torch.Tensor, values: torch.Tensor, count_device: torch.Tensor, start_bit=0,
                             end_bit=None):
  """radix_sort_pairs of the first ``count_device[0]`` pairs of capacity-sized buffers, the count (int32) still on
  the device: nothing is read back; output rows past the count are uninitialised (extension, not in the reference)."""
  check_cuda("keys", keys)
  check_cuda("values", values)
  check_cuda("count", count_device)
  assert keys.dtype in _KEY_BYTES and values.dtype == torch.int32 and count_device.dtype == torch.int32
  assert keys.shape == values.shape and keys.dim() == 1
  kb = _KEY_BYTES[keys.dtype]
  if end_bit is None or end_bit < 0:
    end_bit = kb * 8
  n = keys.shape[0]
  keys_out, values_out = torch.empty_like(keys), torch.empty_like(values)
  if n == 0:
    return keys_out, values_out
  lib = N.lib()
  ws = N.workspace(lib.gs_radix_sort_pairs_workspace_bytes(n, kb, start_bit, end_bit), keys.device)
  N.call("gs_radix_sort_pairs_counted", ctypes.c_int64(n), N.ptr(count_device), ctypes.c_int32(kb),
         N.ptr(keys.contiguous()), N.ptr(values.contiguous()), N.ptr(keys_out), N.ptr(values_out),
         ctypes.c_int32(start_bit), ctypes.c_int32(end_bit), N.ptr(ws), ctypes.c_size_t(ws.numel()),
         N.stream_ptr(keys.device))
  return keys_out, values_out


def radix_argsort(keys: torch.Tensor):
  idx = torch.arange(keys.shape[0], dtype=torch.int32, device=keys.device)
  _, idx = radix_sort_pairs(keys, idx)
  return idx


def segmented_sort_pairs(keys: torch.Tensor, values: torch.Tensor, start_offset: torch.Tensor, end_offset: torch.Tensor):
  """Sort (key, value) pairs by ascending signed key inside every segment [start_offset[i], end_offset[i])
  (cuda_lib/segmented_sort_pairs.cu:9-73: cub::DeviceSegmentedSort::SortPairs; int32 or int16 keys, int32 values, int64
  offsets; segments must not overlap).  Rows outside every segment are copied through.  Not on the render path
  (SURVEY.md K12): composed from the onesweep sort — a stable sort of all rows by key, then a stable sort by segment
  number, then a scatter of every segment to its own offsets — rather than a dedicated kernel."""
  check_cuda("keys", keys)
  check_cuda("values", values)
  assert keys.dim() == 1 and values.dim() == 1 and keys.shape == values.shape, "keys and values must be 1D and equal in size"
  assert start_offset.dim() == 1 and start_offset.shape == end_offset.shape, "offsets must be 1D and equal in size"
  assert start_offset.dtype == torch.int64 and end_offset.dtype == torch.int64, "start_offset/end_offset must be int64"
  assert keys.dtype in (torch.int32, torch.int16) and values.dtype == torch.int32, "Not yet implemented for data type."
  n, nseg = keys.shape[0], start_offset.shape[0]
  keys_out, values_out = keys.clone(), values.clone()
  if n == 0 or nseg == 0:
    return keys_out, values_out
  dev = keys.device
  lengths = (end_offset - start_offset).clamp_min(0)
  # segment number of every row (nseg = outside): +1 / -1 marks at the segment borders, prefix sum
  marks = torch.zeros((n + 1,), dtype=torch.int64, device=dev)
  ids = torch.arange(1, nseg + 1, device=dev)
  live = lengths > 0
  marks.index_add_(0, start_offset[live], ids[live])
  marks.index_add_(0, end_offset[live], -ids[live])
  seg = torch.cumsum(marks[:n], 0) - 1
  seg = torch.where(seg < 0, torch.full_like(seg, nseg), seg).to(torch.int32)
  # pass 1: all rows by key (sign bit flipped: signed order on an unsigned radix sort); pass 2: stable by segment
  flipped = (keys.to(torch.int32) ^ torch.tensor(-2 ** 31, dtype=torch.int32, device=dev)).contiguous()
  rows = torch.arange(n, dtype=torch.int32, device=dev)
  _, order = radix_sort_pairs(flipped, rows, 0, 32)
  seg_sorted = seg[order.long()].contiguous()
  _, order = radix_sort_pairs(seg_sorted, order.contiguous(), 0, max(1, int(nseg).bit_length()))
  order = order.long()
  # grouped row j of segment i goes to start_offset[i] + (j - first grouped row of segment i)
  first = torch.cumsum(lengths, 0) - lengths
  inside = int(lengths.sum().item())
  seg_of = seg[order[:inside]].long()
  dest = start_offset[seg_of] + (torch.arange(inside, device=dev) - first[seg_of])
  keys_out[dest] = keys[order[:inside]]
  values_out[dest] = values[order[:inside]]
  return keys_out, values_out


__all__ = ["full_cumsum", "full_cumsum_device", "radix_sort_pairs", "radix_sort_pairs_counted", "radix_argsort",
           "segmented_sort_pairs"]
