mkdir -p gpurun_out
python -m pytest tests/test_gpu_rasterizer.py tests/test_golden.py tests/test_parameter_class.py tests/test_gpu_renderer.py -m gpu -q 2>&1 | tail -15
python benchmarks/configs.py --only c1 --steps 20 2>&1 | cut -c1-700
