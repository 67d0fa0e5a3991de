set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
nproc
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo smoke rc=$?
tail -3 gpurun_out/r2a_smoke.log
rm -f gpurun_out/parity_at_size.jsonl
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_at_size.py > gpurun_out/r2a_pytest.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/r2a_pytest.log
python -m pytest tests/test_gpu_parity_at_size.py -q -s > gpurun_out/r2a_parity.log 2>&1; echo parity rc=$?
tail -30 gpurun_out/r2a_parity.log | cut -c1-1500
python benchmarks/fast_math_cost.py > gpurun_out/r2a_fastmath.log 2>&1; echo fastmath rc=$?
cat gpurun_out/r2a_fastmath.log | cut -c1-1500
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.log 2>&1; echo bench rc=$?
tail -2 gpurun_out/r2a_bench.log | cut -c1-3000
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.log 2>&1; echo ref rc=$?
tail -1 gpurun_out/r2a_ref.log | cut -c1-600
