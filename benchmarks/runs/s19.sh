mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s19_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/s19_smoke.log | cut -c1-200
python -m pytest tests -m gpu -q > gpurun_out/s19_pytest.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/s19_pytest.log | cut -c1-300
python benchmarks/variants.py --variants 0,1 --scene c3 > gpurun_out/s19_c3_pair.log 2>&1; echo rc=$?
cut -c1-420 gpurun_out/s19_c3_pair.log
