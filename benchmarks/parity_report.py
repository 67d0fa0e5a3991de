#!/usr/bin/env python
"""PARITY.md from gpurun_out/parity_at_size.jsonl (written by tests/test_gpu_parity_at_size.py on the B200 box) plus
the other measured parity figures handed in as JSON lines (benchmarks/fast_math_cost.py output).

  python benchmarks/parity_report.py [--fast-math gpurun_out/r2a_fastmath.log] > PARITY.md
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def fmt(x):
  if x is None:
    return "—"
  if isinstance(x, float):
    return "0" if x == 0 else f"{x:.1e}"
  return str(x)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--rows", default=str(ROOT / "gpurun_out" / "parity_at_size.jsonl"))
  ap.add_argument("--fast-math", default=None)
  ap.add_argument("--run", default="", help="label of the GPU run the numbers come from")
  args = ap.parse_args()
  rows = {}
  for line in Path(args.rows).read_text().splitlines():
    if line.startswith("{"):
      r = json.loads(line)
      rows[r["config"]] = r   # the last run of a configuration wins
  order = [c for c in ("bench", "c2", "c3", "c4", "c5") if c in rows]
  out = []
  out.append("# PARITY — measured on one B200, CUDA path against the oracle at BASELINE.json's full sizes\n")
  out.append(f"Source: `tests/test_gpu_parity_at_size.py` ({args.run}); every number is a relative L2 error "
             "`|cuda - oracle| / |oracle|` unless it says bit-exact.  Tolerances (BASELINE.json north_star): tile maps and "
             "sorted orderings bit-exact; images / depths / features 1e-5; gradients 1e-4.  The oracle (`oracle/`) is pinned "
             "against the reference's own code by `tests/test_golden.py` (see DESIGN.md §4).\n")
  out.append("## Sizes\n")
  out.append("| config | gaussians N | image | F | visible V | overlaps K | K / tile mean | K / tile max | stats | oracle + CUDA seconds |")
  out.append("|---|---:|---|---:|---:|---:|---:|---:|---|---:|")
  for c in order:
    r = rows[c]
    out.append(f"| {c} | {r['N']:,} | {r['image_size'][0]}x{r['image_size'][1]} | {r['F']} | {r['V']:,} | {r['K']:,} | "
               f"{r['K_per_tile_mean']:.0f} | {r['K_per_tile_max']} | {'on' if r['stats'] else 'off'} | {r['seconds']} |")
  out.append("\n`bench` = the workload the metric is quoted on (bench.py); c2..c5 = BASELINE.json configs 2..5.  "
             "Config 1 (2D fit, 20 k gaussians) is covered at full size by `tests/test_gpu_rasterizer.py` / "
             "`tests/test_parameter_class.py`.\n")
  out.append("## Stage by stage\n")
  out.append("| config | visible set, packed gaussians, depths | tile map (overlap_to_point, tile_ranges) | SH colours | image | image weight | "
             "d/d gaussians2d | d/d features | visibility | point heuristic | render_gaussians vs staged |")
  out.append("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
  for c in order:
    r = rows[c]
    out.append(f"| {c} | bit-exact | bit-exact | {fmt(r.get('sh_colours'))} | {fmt(r['image'])} | {fmt(r['image_weight'])} | "
               f"{fmt(r['grad_gaussians2d'])} | {fmt(r['grad_features'])} | {fmt(r.get('visibility'))} | "
               f"{fmt(r.get('point_heuristic'))} | {fmt(r['render_gaussians_vs_staged'])} |")
  out.append("\n(The test asserts the bit-exact columns with `torch.equal` on the int32 views; a row only exists if they held.)\n")
  out.append("## 3D parameter gradients (projection backward), f32\n")
  out.append("CUDA `project_bwd_kernel` against `oracle.projection_backward<float>` — the same reverse sweep in plain IEEE "
             "operations, pinned in f64 against the reference's torch_lib autograd — fed the same upstream gradients.  "
             "\"conditioned\" = gaussians whose projected covariance has sqrt(gap)/trace and |n|/trace above 0.02 "
             "(the two quantities the eigen decomposition divides by, generic.py:216-230); \"all\" includes the rest, where "
             "ANY f32 evaluation is off by O(1/cond) (the f32 restatement itself differs from its f64 instantiation by "
             "1e-2 .. 1e-1 there).\n")
  out.append("| config | conditioned fraction | position | log_scaling | rotation | alpha_logit | all: position | all: log_scaling | all: rotation |")
  out.append("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
  for c in order:
    r = rows[c]
    out.append(f"| {c} | {r['cond_fraction']:.4f} | {fmt(r['grad3d_position'])} | {fmt(r['grad3d_log_scaling'])} | "
               f"{fmt(r['grad3d_rotation'])} | {fmt(r['grad3d_alpha_logit'])} | {fmt(r['grad3d_position_all'])} | "
               f"{fmt(r['grad3d_log_scaling_all'])} | {fmt(r['grad3d_rotation_all'])} |")
  if args.fast_math:
    fm = [json.loads(l) for l in Path(args.fast_math).read_text().splitlines() if l.startswith("{")]
    out.append("\n## What `--use_fast_math` on point_kernels.cu costs (benchmarks/fast_math_cost.py, bench scene)\n")
    out.append("| build | gs_project_bwd ms | log_scaling (conditioned, vs f32 / vs f64) | rotation (conditioned, vs f32 / vs f64) | "
               "position (conditioned) | log_scaling (all, vs f64) | rotation (all, vs f64) |")
    out.append("|---|---:|---|---|---:|---:|---:|")
    for r in fm:
      out.append(f"| {r['build']} | {r['ms']['gs_project_bwd']} | {fmt(r['log_scaling']['conditioned_vs_f32'])} / "
                 f"{fmt(r['log_scaling']['conditioned_vs_f64'])} | {fmt(r['rotation']['conditioned_vs_f32'])} / "
                 f"{fmt(r['rotation']['conditioned_vs_f64'])} | {fmt(r['position']['conditioned_vs_f64'])} | "
                 f"{fmt(r['log_scaling']['all_vs_f64'])} | {fmt(r['rotation']['all_vs_f64'])} |")
    out.append("\nThe MUFU-based division / sqrt / exp change nothing that the tolerance can see (1.8e-5 against 1.75e-5 on "
               "the conditioned gaussians) and save 0.033 ms per frame (stand-alone calls with plain gradient writes: 0.146 "
               "against 0.179 ms); the large unconditioned numbers are the same with and without the flag — they are f32 "
               "cancellation in the eigen decomposition, not the approximations.  The flag stays.\n")
  extra = ROOT / "benchmarks" / "parity_extra.md"   # sections measured by other tests (static path, multi-GPU sums)
  if extra.exists():
    out.append("\n" + extra.read_text().rstrip("\n"))
  print("\n".join(out))


if __name__ == "__main__":
  main()
