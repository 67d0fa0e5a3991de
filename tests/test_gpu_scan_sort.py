"""cuda_lib surface on the GPU: full_cumsum and radix_sort_pairs, bit-exact against torch / the oracle.
The reference has no test at this boundary (parity unpinned there); integer results are
implementation independent given the stability contract."""
import numpy as np
import pytest
import torch

import oracle
from taichi_gaussian_rasterizer_b200 import cuda_lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 31, 2048, 2049, 100_003, 3_000_001])
@pytest.mark.parametrize("dtype", [torch.int32, torch.int64])
def test_full_cumsum(cuda_device, n, dtype):
  torch.manual_seed(n)
  x = torch.randint(0, 50, (n,), dtype=dtype, device=cuda_device)
  out, total = cuda_lib.full_cumsum(x)
  assert out.shape == (n + 1,) and out.dtype == dtype
  ref = torch.zeros(n + 1, dtype=torch.int64)
  ref[1:] = torch.cumsum(x.cpu().long(), 0)
  assert total == int(ref[-1])
  assert torch.equal(out.cpu().long(), ref)


@pytest.mark.parametrize("n", [1, 2, 255, 4096, 4097, 50_000, 2_000_003])
@pytest.mark.parametrize("key_dtype,bits", [(torch.int64, (0, 48)), (torch.int64, (0, 46)), (torch.int64, (0, 64)),
                                            (torch.int32, (0, 32)), (torch.int32, (0, 30)), (torch.int64, (32, 46))])
def test_radix_sort_pairs_is_a_stable_sort(cuda_device, n, key_dtype, bits):
  torch.manual_seed(n + bits[1])
  start, end = bits
  hi = 1 << min(end, 62 if key_dtype == torch.int64 else 31)
  # few distinct values in the sorted bits -> many ties -> stability is exercised
  distinct = max(2, min(n // 8 + 2, hi))
  keys = torch.randint(0, distinct, (n,), dtype=torch.int64) * (hi // distinct)
  if key_dtype == torch.int64 and end < 64:
    keys = keys | (torch.randint(0, 1 << 10, (n,), dtype=torch.int64) << min(end + 2, 52))  # noise above end_bit
  values = torch.arange(n, dtype=torch.int32)
  k_out, v_out = cuda_lib.radix_sort_pairs(keys.to(key_dtype).to(cuda_device), values.to(cuda_device), start, end)
  mask = ((1 << (end - start)) - 1)
  sort_bits = (keys.numpy().astype(np.int64).view(np.uint64) >> np.uint64(start)) & np.uint64(mask)
  order = np.argsort(sort_bits, kind="stable")
  assert np.array_equal(v_out.cpu().numpy(), values.numpy()[order])
  assert np.array_equal(k_out.cpu().numpy().astype(np.int64), keys.to(key_dtype).numpy().astype(np.int64)[order])
  # the oracle's LSD sort agrees too
  ko, vo = oracle.radix_sort_pairs(keys.to(key_dtype).long() & (0xFFFFFFFF if key_dtype == torch.int32 else -1),
                                   values, start, end)
  assert torch.equal(vo, v_out.cpu())


def test_sort_preserves_inputs_and_runs_on_current_stream(cuda_device):
  keys = torch.randint(0, 1 << 40, (100_000,), dtype=torch.int64, device=cuda_device)
  values = torch.arange(100_000, dtype=torch.int32, device=cuda_device)
  k0, v0 = keys.clone(), values.clone()
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    k_out, v_out = cuda_lib.radix_sort_pairs(keys, values, 0, 48)
  s.synchronize()
  assert torch.equal(keys, k0) and torch.equal(values, v0)
  assert torch.equal(k_out & ((1 << 48) - 1), torch.sort(k0 & ((1 << 48) - 1), stable=True).values)


@pytest.mark.parametrize("key_dtype", [torch.int32, torch.int16])
def test_segmented_sort_pairs(cuda_device, key_dtype):
  """cuda_lib.segmented_sort_pairs (cuda_lib/segmented_sort_pairs.cu:9-73, the demo of cuda_lib/__init__.py:47-60):
  ascending signed keys inside every segment, rows outside the segments untouched, values follow their keys."""
  from taichi_gaussian_rasterizer_b200.cuda_lib import segmented_sort_pairs
  torch.manual_seed(4)
  n = 5000
  hi = 2 ** 15 if key_dtype == torch.int16 else 2 ** 31
  keys = torch.randint(-hi, hi, (n,), dtype=torch.int64).to(key_dtype)
  values = torch.arange(n, dtype=torch.int32)
  cuts = torch.sort(torch.randperm(n - 1)[:40] + 1).values
  bounds = torch.cat([torch.tensor([0]), cuts, torch.tensor([n])])
  start, end = bounds[:-1].clone(), bounds[1:].clone()
  start, end = torch.cat([start[:10], start[12:]]), torch.cat([end[:10], end[12:]])   # a gap: two segments left out
  end[3] = start[3]                                                                  # and an empty segment
  ko, vo = segmented_sort_pairs(keys.to(cuda_device), values.to(cuda_device), start.to(cuda_device), end.to(cuda_device))
  ko, vo = ko.cpu(), vo.cpu()
  covered = torch.zeros(n, dtype=torch.bool)
  for s, e in zip(start.tolist(), end.tolist()):
    if e > s:
      covered[s:e] = True
      ref = torch.sort(keys[s:e].long()).values
      assert torch.equal(ko[s:e].long(), ref)
      assert torch.equal(keys[vo[s:e].long()], ko[s:e])                 # values follow their keys
      assert torch.equal(torch.sort(vo[s:e]).values, values[s:e])       # and stay inside their segment
  assert torch.equal(ko[~covered], keys[~covered]) and torch.equal(vo[~covered], values[~covered])
  e0, e1 = segmented_sort_pairs(keys[:0].to(cuda_device), values[:0].to(cuda_device), start[:0].to(cuda_device),
                                end[:0].to(cuda_device))
  assert e0.shape == (0,) and e1.shape == (0,)
