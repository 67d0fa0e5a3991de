python -m pytest tests/test_gpu_rasterizer.py tests/test_golden.py -m gpu -q -k "wide or golden or fast" 2>&1 | tail -15
python benchmarks/variants.py --variants 0,2 --scene c4 --rounds 3 --iters 5 2>&1 | cut -c1-700
