mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/s10_pytest.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/s10_pytest.log | cut -c1-300
python benchmarks/variants.py --variants 0 --scene c3 > gpurun_out/s10_c3_stats.log 2>&1; echo rc=$?
cut -c1-500 gpurun_out/s10_c3_stats.log
