set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rasterizer.py tests/test_gpu_renderer.py tests/test_golden.py -m gpu -q -x > gpurun_out/s1_pytest.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/s1_pytest.log
python benchmarks/variants.py --variants 0,1 --scene bench > gpurun_out/s1_variants.log 2>&1; echo variants rc=$?
tail -4 gpurun_out/s1_variants.log | cut -c1-1200
python benchmarks/configs.py --only c3,c5 > gpurun_out/s1_configs.log 2>&1; echo configs rc=$?
cut -c1-900 gpurun_out/s1_configs.log
python bench.py --no-cpu-baseline --no-configs --no-stock --steps 5 --warmup 3 > gpurun_out/s1_bench.log 2>&1; echo bench rc=$?
tail -1 gpurun_out/s1_bench.log | cut -c1-400
