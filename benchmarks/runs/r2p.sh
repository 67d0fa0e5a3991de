mkdir -p gpurun_out
for st in 1 2 3; do
python bench.py --steps 8 --warmup 3 --no-configs --no-cpu-baseline --no-stock --streams $st > gpurun_out/r2p_s$st.log 2>&1; echo rc=$?
tail -1 gpurun_out/r2p_s$st.log | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('streams', d['config']['view_streams'], 'ms/frame', round(d['ms_per_frame'],4), 'e2e', round(d['e2e']['ms_per_frame'],4))"
done
python -m pytest tests/test_gpu_renderer.py -m gpu -q 2>&1 | tail -2
