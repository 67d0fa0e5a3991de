mkdir -p gpurun_out
python -m pytest tests/test_gpu_static_render.py -m gpu -q > gpurun_out/s17_pytest.log 2>&1; echo pytest rc=$?
tail -25 gpurun_out/s17_pytest.log | cut -c1-300
