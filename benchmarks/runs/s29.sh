mkdir -p gpurun_out
python -m pytest tests/test_gpu_rasterizer.py tests/test_golden.py -m gpu -q -x > gpurun_out/s29_pytest.log 2>&1; echo pytest rc=$?
tail -2 gpurun_out/s29_pytest.log
python bench.py --no-cpu-baseline --no-configs --no-stock --steps 10 > gpurun_out/s29_bench.log 2> gpurun_out/s29_bench.err; echo bench rc=$?
python - <<P
import json
for l in open('gpurun_out/s29_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_frame'],4), round(d['e2e']['ms_per_frame'],4), round(d['eager_ms_per_frame'],4), d['stage_ms_per_frame']['gs_raster_fwd'], d['stage_ms_per_frame']['gs_raster_bwd'])
P
python benchmarks/variants.py --variants 0 --scene c4 --rounds 3 --iters 5 > gpurun_out/s29_c4.log 2>&1; cut -c1-250 gpurun_out/s29_c4.log
python benchmarks/variants.py --variants 0 --scene c3 --rounds 3 --iters 5 > gpurun_out/s29_c3.log 2>&1; cut -c1-250 gpurun_out/s29_c3.log
