"""Scene IO in the gaussian-splatting ``point_cloud.ply`` format — what the reference's published benchmark protocol
loads (BENCHMARK.md:19-23: reconstructions of the official gaussian-splatting implementation, read by the external
splat-viewer and converted "to the most convenient form" before timing).  The reference repository itself carries no
reader; this one goes straight to the layout the render path takes (``Gaussians3D``).

File layout (binary little endian, one ``vertex`` element, float32 properties, any order):
  x y z | nx ny nz (ignored) | f_dc_0..2 | f_rest_0..3K-1 (channel major: all K of red, then green, then blue) |
  opacity (logit) | scale_0..2 (log) | rot_0..3 (quaternion w x y z)
and ``Gaussians3D`` wants position, log_scaling, rotation in x y z w order (taichi_lib/generic.py:418-427),
alpha_logit (N, 1) and feature (N, 3, (degree + 1)^2) with the DC term first.
"""
from pathlib import Path
from typing import Dict, List, Tuple, Union

import numpy as np
import torch

from ..data_types import Gaussians3D

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def _read_header(f) -> Tuple[str, int, List[Tuple[str, str]]]:
  if f.readline().strip() != b"ply":
    raise ValueError("not a PLY file")
  fmt, count, props, in_vertex = None, None, [], False
  while True:
    line = f.readline()
    if not line:
      raise ValueError("PLY header without end_header")
    words = line.decode("ascii").split()
    if not words or words[0] == "comment":
      continue
    if words[0] == "format":
      fmt = words[1]
    elif words[0] == "element":
      in_vertex = words[1] == "vertex"
      if in_vertex:
        count = int(words[2])
      elif count is None:
        raise ValueError("PLY elements before 'vertex' are not supported")
    elif words[0] == "property" and in_vertex:
      if words[1] == "list":
        raise ValueError("list properties in the vertex element are not supported")
      props.append((words[2], _PLY_TYPES[words[1]]))
    elif words[0] == "end_header":
      break
  if fmt not in ("binary_little_endian", "binary_big_endian", "ascii") or count is None:
    raise ValueError(f"unsupported PLY header (format {fmt}, vertex count {count})")
  return fmt, count, props


def read_ply_vertices(path: Union[str, Path]) -> Dict[str, np.ndarray]:
  """All vertex properties of a PLY file as a dict of 1-D arrays."""
  with open(path, "rb") as f:
    fmt, count, props = _read_header(f)
    if fmt == "ascii":
      table = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2)
      return {name: table[:, i].astype(t) for i, (name, t) in enumerate(props)}
    order = "<" if fmt == "binary_little_endian" else ">"
    dtype = np.dtype([(name, order + t) for name, t in props])
    data = np.frombuffer(f.read(count * dtype.itemsize), dtype=dtype, count=count)
  return {name: data[name] for name, _ in props}


def load_ply(path: Union[str, Path], device=None, max_sh_degree: int = None) -> Gaussians3D:
  """A gaussian-splatting point cloud as ``Gaussians3D`` with spherical-harmonics features (N, 3, (degree + 1)^2)."""
  v = read_ply_vertices(path)
  n = v["x"].shape[0]

  def cols(prefix, k):
    return np.stack([v[f"{prefix}{i}"].astype(np.float32) for i in range(k)], axis=1) if k else np.zeros((n, 0), np.float32)

  n_rest = sum(1 for name in v if name.startswith("f_rest_"))
  if n_rest % 3 != 0 or int(round((n_rest // 3 + 1) ** 0.5)) ** 2 != n_rest // 3 + 1:
    raise ValueError(f"{n_rest} f_rest_* properties are not 3 * ((degree + 1)^2 - 1)")
  dc = cols("f_dc_", 3)[:, :, None]                                   # (N, 3, 1)
  rest = cols("f_rest_", n_rest).reshape(n, 3, n_rest // 3)           # channel major
  feature = np.concatenate([dc, rest], axis=2)
  if max_sh_degree is not None:
    feature = feature[:, :, :(max_sh_degree + 1) ** 2]
  wxyz = cols("rot_", 4)
  g = Gaussians3D(position=torch.from_numpy(np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float32)),
                  log_scaling=torch.from_numpy(cols("scale_", 3)),
                  rotation=torch.from_numpy(np.ascontiguousarray(wxyz[:, [1, 2, 3, 0]])),
                  alpha_logit=torch.from_numpy(v["opacity"].astype(np.float32).reshape(n, 1)),
                  feature=torch.from_numpy(np.ascontiguousarray(feature)), batch_size=(n,))
  return g.to(device=device) if device is not None else g


def save_ply(gaussians: Gaussians3D, path: Union[str, Path]):
  """Writes ``gaussians`` (SH features (N, 3, K) or plain RGB (N, 3)) as a gaussian-splatting point cloud."""
  g = gaussians.detach().cpu()
  n = g.position.shape[0]
  feature = g.feature if g.feature.ndim == 3 else g.feature[:, :, None]
  assert feature.shape[1] == 3, f"PLY scenes carry 3 colour channels, got {tuple(feature.shape)}"
  k = feature.shape[2]
  names = (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] +
           [f"f_rest_{i}" for i in range(3 * (k - 1))] + ["opacity"] + [f"scale_{i}" for i in range(3)] +
           [f"rot_{i}" for i in range(4)])
  table = np.concatenate([
    g.position.numpy(), np.zeros((n, 3), np.float32), feature[:, :, 0].numpy(),
    feature[:, :, 1:].reshape(n, 3 * (k - 1)).numpy(), g.alpha_logit.reshape(n, 1).numpy(), g.log_scaling.numpy(),
    g.rotation[:, [3, 0, 1, 2]].numpy()], axis=1).astype("<f4")
  header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n" + \
           "".join(f"property float {name}\n" for name in names) + "end_header\n"
  with open(path, "wb") as f:
    f.write(header.encode("ascii"))
    f.write(np.ascontiguousarray(table).tobytes())


__all__ = ["load_ply", "save_ply", "read_ply_vertices"]
