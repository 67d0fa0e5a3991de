set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rasterizer.py tests/test_golden.py -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/r2c_pytest.log
python benchmarks/variants.py --variants 0,1 --scene bench > gpurun_out/r2c_variants.log 2>&1; echo rc=$?
python benchmarks/variants.py --variants 0,1 --scene c3 >> gpurun_out/r2c_variants.log 2>&1; echo rc=$?
python benchmarks/variants.py --variants 0,1 --scene c2 >> gpurun_out/r2c_variants.log 2>&1; echo rc=$?
cat gpurun_out/r2c_variants.log
