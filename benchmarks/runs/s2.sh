set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_static_render.py tests/test_gpu_rasterizer.py tests/test_gpu_renderer.py tests/test_golden.py -m gpu -q > gpurun_out/s2_pytest.log 2>&1; echo pytest rc=$?
tail -30 gpurun_out/s2_pytest.log | cut -c1-400
python benchmarks/variants.py --variants 0 --scene c3 > gpurun_out/s2_c3_stats.log 2>&1; echo rc=$?
python benchmarks/variants.py --variants 0 --scene c3 --no-stats > gpurun_out/s2_c3_nostats.log 2>&1; echo rc=$?
cut -c1-500 gpurun_out/s2_c3_stats.log gpurun_out/s2_c3_nostats.log
python bench.py --no-cpu-baseline --no-configs --no-stock --steps 5 --warmup 3 > gpurun_out/s2_bench.log 2>&1; echo bench rc=$?
tail -1 gpurun_out/s2_bench.log | cut -c1-1800
tail -5 gpurun_out/s2_bench.log | grep -i "error\|Traceback" | head
