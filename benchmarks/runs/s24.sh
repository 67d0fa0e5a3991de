mkdir -p gpurun_out
python -m pytest tests/test_gpu_rasterizer.py tests/test_golden.py -m gpu -q -x > gpurun_out/s24_pytest.log 2>&1; echo pytest rc=$?
tail -2 gpurun_out/s24_pytest.log
python benchmarks/variants.py --variants 0 --scene c4 --rounds 5 --iters 5 > gpurun_out/s24_c4.log 2>&1; echo rc=$?
cut -c1-300 gpurun_out/s24_c4.log
