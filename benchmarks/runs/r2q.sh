python -m pytest tests/test_gpu_renderer.py -m gpu -q -k "streams" 2>&1 | tail -6
python bench.py --steps 8 --warmup 3 --no-configs --no-cpu-baseline --no-stock 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('streams', d['config']['view_streams'], 'ms/frame', round(d['ms_per_frame'],4), 'e2e', round(d['e2e']['ms_per_frame'],4))"
