import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
  sys.path.insert(0, str(ROOT))


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
  """The oracle and the C-ABI library are built in-tree; build them if a fresh checkout lacks them."""
  import oracle
  oracle.build()
  from taichi_gaussian_rasterizer_b200.csrc import build as build_ext
  build_ext.build()
  yield


@pytest.fixture(scope="session")
def cuda_device():
  import torch
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  return torch.device("cuda:0")
