#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv profiles/r01_launches.md [--frames N]
  python profiles/summarize.py report   gpurun_out/prof_raster.ncu-rep profiles/r01_raster_ncu.md
  python profiles/summarize.py hot      gpurun_out/prof_raster.ncu-rep raster_bwd profiles/r01_raster_bwd_hot.md

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (times are cold-cache and
serialised under ncu: read the SHARES).  `report` pulls the per launch metrics the roofline uses out of a
`--set full` capture (dram bytes, duration, registers, occupancy, pipe utilisation).  `hot` lists the SASS
instructions that execute most, with the stall samples (needs -lineinfo / --import-source on).
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
  "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
  "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
  "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
  "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
  "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
  "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
  "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def ncu_csv(rep, page, extra=()):
  out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], stdout=subprocess.PIPE,
                       stderr=subprocess.DEVNULL, text=True).stdout
  return list(csv.reader(io.StringIO(out)))


def short(name, n=90):
  name = name.replace("void ", "").replace("gs::", "")
  return name if len(name) <= n else name[:n - 1] + "…"


def launches(src, dst, frames=None):
  rows = list(csv.reader(open(src)))
  h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
  hd = rows[h]
  ki, vi, ui = hd.index("Kernel Name"), hd.index("Metric Value"), hd.index("Metric Unit")
  agg = collections.OrderedDict()
  for r in rows[h + 1:]:
    if len(r) <= vi:
      continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
  tot = sum(v[1] for v in agg.values())
  ours = sum(v[1] for k, v in agg.items() if "gs::" in k)
  with open(dst, "w") as f:
    f.write(f"# launch list summary of `{src}`\n\n")
    f.write("ncu `--metrics gpu__time_duration.sum --clock-control none`: per launch device time, cold cache and "
            "serialised, so compare the SHARES with the live CUDA-event numbers of bench.py, not the absolutes.\n\n")
    f.write(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches; kernels of libgsplat_b200.so "
            f"(gs::*): {ours:.3f} ms = {100 * ours / tot:.1f} %; the rest is torch glue (loss, grad accumulation, "
            f"fills, camera inverse).\n\n")
    f.write("| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      if v[1] / tot < 0.0005:
        continue
      f.write(f"| `{short(k)}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0]:.4f} |\n")
  print(f"wrote {dst}")


def report(rep, dst):
  rows = ncu_csv(rep, "raw")
  hd, units = rows[0], rows[1]
  with open(dst, "w") as f:
    f.write(f"# `ncu --set full --clock-control none` capture `{rep}`\n\nper launch values.\n")
    for r in rows[2:]:
      if len(r) < len(hd):
        continue
      f.write(f"\n## `{short(r[hd.index('Kernel Name')], 160)}`\n\n| metric | value | unit |\n|---|---:|---|\n")
      for k in KEYS:
        if k in hd:
          f.write(f"| {k} | {r[hd.index(k)]} | {units[hd.index(k)]} |\n")
      try:
        rd = float(r[hd.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[
          units[hd.index("dram__bytes_read.sum")]]
        wr = float(r[hd.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[
          units[hd.index("dram__bytes_write.sum")]]
        f.write(f"| **traffic = dram read + write** | {(rd + wr) / 1e6:.1f} | MB |\n")
      except Exception:
        pass
  print(f"wrote {dst}")


def hot(rep, pattern, dst, top=60):
  rows = ncu_csv(rep, "source", ["--kernel-name", f"regex:{pattern}"])
  hd = rows[1]
  ie, isamp, isrc, ia = (hd.index("Instructions Executed"), hd.index("# Samples"), hd.index("Source"),
                         hd.index("Avg. Predicated-On Threads Executed"))
  blk = []
  for r in rows[2:]:
    if r and r[0] == "Kernel Name":
      break
    if len(r) > ie:
      blk.append(r)
  tot = sum(int(r[ie]) for r in blk)
  ts = max(sum(int(r[isamp]) for r in blk), 1)
  ops = collections.Counter()
  for r in blk:
    ops[r[isrc].split()[0] if not r[isrc].lstrip().startswith("@") else r[isrc].split()[1]] += int(r[ie])
  with open(dst, "w") as f:
    f.write(f"# hottest SASS of `{rows[0][1] if len(rows[0]) > 1 else pattern}`\n\n")
    f.write(f"{len(blk)} SASS instructions, {tot} warp instructions executed, {ts} stall samples.\n\n")
    f.write("## opcode mix (warp instructions executed)\n\n| opcode | share |\n|---|---:|\n")
    for op, n in ops.most_common(16):
      f.write(f"| {op} | {100 * n / tot:.1f} % |\n")
    f.write(f"\n## top {top} by stall samples\n\n| # | executed (M) | exec share | sample share | active thr | SASS |\n"
            "|---:|---:|---:|---:|---:|---|\n")
    order = sorted(range(len(blk)), key=lambda i: -int(blk[i][isamp]))[:top]
    for i in order:
      r = blk[i]
      f.write(f"| {i} | {int(r[ie]) / 1e6:.2f} | {100 * int(r[ie]) / tot:.2f} % | {100 * int(r[isamp]) / ts:.2f} % | "
              f"{float(r[ia]):.1f} | `{r[isrc].strip()[:90]}` |\n")
  print(f"wrote {dst}")


if __name__ == "__main__":
  cmd = sys.argv[1]
  if cmd == "launches":
    launches(sys.argv[2], sys.argv[3])
  elif cmd == "report":
    report(sys.argv[2], sys.argv[3])
  elif cmd == "hot":
    hot(sys.argv[2], sys.argv[3], sys.argv[4])
  else:
    raise SystemExit(__doc__)
