set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_at_size.jsonl
python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_at_size.py > gpurun_out/r2b_pytest.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/r2b_pytest.log
python -m pytest tests/test_gpu_parity_at_size.py -q -s > gpurun_out/r2b_parity.log 2>&1; echo parity rc=$?
grep '^{' gpurun_out/r2b_parity.log | cut -c1-1200
tail -5 gpurun_out/r2b_parity.log | cut -c1-600
