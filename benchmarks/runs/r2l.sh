python -m pytest tests/test_gpu_tile_mapper.py -m gpu -q 2>&1 | tail -15
python benchmarks/configs.py --only c1,c1_graph --steps 50 2>&1 | cut -c1-600
