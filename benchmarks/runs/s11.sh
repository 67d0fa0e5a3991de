mkdir -p gpurun_out
python -m pytest tests/test_gpu_rasterizer.py tests/test_golden.py -m gpu -q -x -k "wide or F34 or 34 or golden or feature" > gpurun_out/s11_pytest.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/s11_pytest.log
python benchmarks/variants.py --variants 0,16,24 --scene c4 --rounds 5 --iters 5 > gpurun_out/s11_c4.log 2>&1; echo rc=$?
cut -c1-600 gpurun_out/s11_c4.log
