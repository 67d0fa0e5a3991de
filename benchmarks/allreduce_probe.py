#!/usr/bin/env python
"""How fast can the step's gradient bucket (708 MB of f32 at 3 M gaussians, SH degree 3) be summed over the ranks of one
NVSwitch box?  Times, device side (max over ranks), NCCL's all-reduce against the in-switch reduction through a
multicast mapping (torch symmetric memory: multimem.ld_reduce / multimem.st) — torch's own kernel and ours
(csrc/multimem_reduce.cu, gs_multimem_all_reduce) — and checks the sums against each other.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/allreduce_probe.py
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def timed(fn, iters, device):
  for _ in range(3):
    fn()
  dist.barrier()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  t = torch.tensor([a.elapsed_time(b) / iters], device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--floats", type=int, default=59 * 3_000_000)
  ap.add_argument("--iters", type=int, default=10)
  args = ap.parse_args()
  rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
  torch.cuda.set_device(local)
  device = torch.device("cuda", local)
  dist.init_process_group("nccl", device_id=device)
  n = args.floats
  out = {"world": world, "mbytes": n * 4 / 1e6}

  torch.manual_seed(rank)
  src = torch.randn(n, device=device)
  ref = src.clone()
  dist.all_reduce(ref)

  buf = src.clone()
  out["nccl_ms"] = timed(lambda: dist.all_reduce(buf), args.iters, device)

  try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(n, dtype=torch.float32, device=device)
    h = symm_mem.rendezvous(t, dist.group.WORLD)
    out["multicast_ptr"] = int(h.multicast_ptr)
    out["signal_pad_size"] = int(h.signal_pad_size)
    group_name = dist.group.WORLD.group_name
    for name in ("multimem_all_reduce_", "two_shot_all_reduce_"):
      op = getattr(torch.ops.symm_mem, name, None)
      if op is None or (name.startswith("multimem") and not h.multicast_ptr):
        continue
      t.copy_(src)
      dist.barrier()
      op(t, "sum", group_name)
      torch.cuda.synchronize()
      out[name + "rel_err"] = float(((t - ref).norm() / ref.norm()).item())
      out[name + "ms"] = timed(lambda: op(t, "sum", group_name), args.iters, device)
    try:
      from taichi_gaussian_rasterizer_b200.distributed import SymmetricBucketReducer
      for blocks in (8, 16, 32, 64, 128):
        red = SymmetricBucketReducer.create(n, device, blocks=blocks)
        assert red is not None, "no multicast mapping"
        red.buffer.copy_(src)
        dist.barrier()
        red.all_reduce()
        torch.cuda.synchronize()
        out[f"ours_b{blocks}_rel_err"] = float(((red.buffer - ref).norm() / ref.norm()).item())
        out[f"ours_b{blocks}_ms"] = timed(red.all_reduce, args.iters, device)
        del red
    except Exception as e:   # noqa: BLE001
      out["ours_error"] = f"{type(e).__name__}: {e}"[:300]
  except Exception as e:   # noqa: BLE001
    out["symm_error"] = f"{type(e).__name__}: {e}"[:300]
  if rank == 0:
    print(json.dumps(out), flush=True)
  dist.destroy_process_group()


if __name__ == "__main__":
  main()
