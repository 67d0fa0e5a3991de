"""How much does a forward early exit buy, and what does it cost?  Renders the bench scene with
forward_exit_transmittance = eps (a warp stops once every pixel's transmittance 1 - W <= eps) and reports the
rasterizer forward kernel time and the deviation from the exact (eps = 0) result.  Outcome on the bench scene
(profiles/r01d_forward_exit.jsonl): the median final transmittance is 2.5e-4, so almost no warp can stop early and
the kernel time does not move; the default stays 0 (exact, the reference has no forward exit: SURVEY Q2)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from taichi_gaussian_rasterizer_b200 import _native
from taichi_gaussian_rasterizer_b200 import RasterConfig, render_gaussians, set_raster_options
W=dict(bench.WORKLOAD)
dev=torch.device('cuda:0')
g_cpu,cams=bench.build_scene(W['num_gaussians'],W['image_size'],W['sh_degree'],W['scale_factor'],W['alpha_range'],W['seed'],1)
g=g_cpu.to(device=dev); cam=cams[0].to(device=dev)
cfg=RasterConfig(compute_visibility=True)
def run(eps):
    set_raster_options(forward_exit_transmittance=eps)
    with torch.no_grad():
        r=render_gaussians(g,cam,cfg,use_sh=True)
    torch.cuda.synchronize()
    timer=_native.set_stage_timer(_native.StageTimer())
    for _ in range(10):
        with torch.no_grad(): r=render_gaussians(g,cam,cfg,use_sh=True)
    torch.cuda.synchronize()
    _native.set_stage_timer(None)
    n,ms=timer.summary()["gs_raster_fwd"]
    return r, ms/n
ref,t0=run(0.0)
for eps in [2**-24, 2**-20, 2**-17, 1e-4]:
    r,t=run(eps)
    rel=lambda x,y: ((x.double()-y.double()).norm()/y.double().norm()).item()
    print(json.dumps(dict(eps=eps, raster_fwd_ms=round(t,4), exact_raster_fwd_ms=round(t0,4), image_rel_l2=rel(r.image,ref.image), weight_rel_l2=rel(r.image_weight,ref.image_weight), vis_rel_l2=rel(r.point_visibility,ref.point_visibility), max_abs=(r.image-ref.image).abs().max().item(), minT=float((1-ref.image_weight).min()), medianT=float((1-ref.image_weight).median()))))
