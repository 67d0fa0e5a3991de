mkdir -p gpurun_out
python -m pytest tests/test_gpu_static_render.py -m gpu -q -x > gpurun_out/s16_pytest.log 2>&1; echo pytest rc=$?
tail -12 gpurun_out/s16_pytest.log | cut -c1-300
python benchmarks/configs.py --only c1_graph,c2_graph > gpurun_out/s16_configs.log 2>&1; cut -c1-200 gpurun_out/s16_configs.log
python bench.py --no-cpu-baseline --no-configs --steps 8 > gpurun_out/s16_bench.log 2> gpurun_out/s16_bench.err; echo bench rc=$?
tail -3 gpurun_out/s16_bench.err | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/s16_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('ms_per_frame','eager_ms_per_frame','without_stale_tail_ms_per_frame','stock_api_ms_per_frame')}, d['e2e']['ms_per_frame'], d['config']['cuda_graph'], d['config']['cuda_graph_error'], d['config']['graph_vs_eager_grad_rel_l2'])
P
