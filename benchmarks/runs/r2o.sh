mkdir -p gpurun_out
rm -f gpurun_out/parity_at_size.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo pytest rc=$?
tail -6 gpurun_out/r2o_pytest.log
