python -m pytest tests/test_gpu_rasterizer.py tests/test_golden.py -m gpu -q -k "wide or golden or feature" 2>&1 | tail -12
python benchmarks/variants.py --variants 0,4 --scene c4 --rounds 3 --iters 5 2>&1 | cut -c1-420
