"""Times the two sorts of the depth-first tile mapper in isolation (CUDA events, L2 flushed between runs):
V depth keys on 32 bits and K tile ids on ceil(log2 T) bits, next to torch.sort on the same keys.

  python benchmarks/sort_bench.py [--v 3000000] [--k 6650000] [--tiles 11008]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from taichi_gaussian_rasterizer_b200 import cuda_lib  # noqa: E402


def timed(fn, flush, reps=20):
  ts = []
  for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1))
  ts.sort()
  return ts[len(ts) // 2]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--v", type=int, default=3_000_000)
  ap.add_argument("--k", type=int, default=6_650_000)
  ap.add_argument("--tiles", type=int, default=11008)
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  torch.manual_seed(0)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  depth = torch.rand(args.v, device=dev)
  dkeys = depth.view(torch.int32)
  vals = torch.arange(args.v, dtype=torch.int32, device=dev)
  bits = max(1, (args.tiles - 1).bit_length())
  tkeys = torch.randint(0, args.tiles, (args.k,), dtype=torch.int32, device=dev)
  tvals = torch.arange(args.k, dtype=torch.int32, device=dev)
  for _ in range(3):
    cuda_lib.radix_sort_pairs(dkeys, vals, 0, 32)
    cuda_lib.radix_sort_pairs(tkeys, tvals, 0, bits)
  k1, v1 = cuda_lib.radix_sort_pairs(dkeys, vals, 0, 32)
  rk, ri = torch.sort(dkeys, stable=True)
  ok_v = bool(torch.equal(k1, rk) and torch.equal(v1.long(), ri))
  k2, v2 = cuda_lib.radix_sort_pairs(tkeys, tvals, 0, bits)
  rk2, ri2 = torch.sort(tkeys, stable=True)
  ok_k = bool(torch.equal(k2, rk2) and torch.equal(v2.long(), ri2))
  out = {
    "v": args.v, "k": args.k, "tile_bits": bits, "correct": ok_v and ok_k,
    "depth_sort_ms": timed(lambda: cuda_lib.radix_sort_pairs(dkeys, vals, 0, 32), flush),
    "tile_sort_ms": timed(lambda: cuda_lib.radix_sort_pairs(tkeys, tvals, 0, bits), flush),
    "torch_sort_depth_ms": timed(lambda: torch.sort(dkeys, stable=True), flush),
    "torch_sort_tile_ms": timed(lambda: torch.sort(tkeys, stable=True), flush),
  }
  out["depth_keys_per_s"] = args.v / (out["depth_sort_ms"] * 1e-3)
  out["tile_keys_per_s"] = args.k / (out["tile_sort_ms"] * 1e-3)
  print(json.dumps(out))


if __name__ == "__main__":
  main()
