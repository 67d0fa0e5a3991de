"""The view-parallel step on real GPUs over NCCL: tests/multi_gpu_worker.py under torchrun, one
process per GPU.  Every rank renders its share of a batch of views into the flat gradient bucket (fused accumulation,
deferred SH gradient, batched SH colours) and the ranks all-reduce once; rank 0 then renders ALL views alone with plain
autograd accumulation and compares (relative L2 < 1e-5).  Skipped on boxes with fewer than two GPUs; the host logic of
the same path is covered on CPU by tests/test_distributed.py (gloo, world size 2)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_all_reduced_gradients_equal_single_gpu_sum(cuda_device, world):
  if torch.cuda.device_count() < world:
    pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tests" / "multi_gpu_worker.py")]
  r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=str(ROOT))
  print(r.stdout[-2000:])
  assert r.returncode == 0, r.stdout[-4000:]
  assert "-> OK" in r.stdout


def test_operators_follow_the_tensors_device(cuda_device):
  """Tensors on cuda:1 while cuda:0 is the current device: the C-ABI launches must run on the tensors' device
  (_native.call's device guard) and give the same image as on cuda:0."""
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  import sys as _sys
  _sys.path.insert(0, str(ROOT / "tests"))
  from taichi_gaussian_rasterizer_b200 import RasterConfig, render_gaussians
  from util import scene3d
  g, cam = scene3d(5, 4000, image_size=(256, 160), scale_factor=0.7, sh_degree=3)
  cfg = RasterConfig(compute_visibility=True)
  images = []
  assert torch.cuda.current_device() == 0
  for index in (0, 1):
    dev = torch.device("cuda", index)
    gd = g.to(device=dev).requires_grad_(True)
    out = render_gaussians(gd, cam.to(device=dev), cfg, use_sh=True)
    out.image.sum().backward()
    assert out.image.device == dev and gd.position.grad.device == dev
    images.append((out.image.detach().cpu(), gd.feature.grad.cpu()))
  assert torch.cuda.current_device() == 0
  assert torch.equal(images[0][0], images[1][0])
  assert torch.allclose(images[0][1], images[1][1], rtol=1e-4, atol=1e-7)
