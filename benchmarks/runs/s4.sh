mkdir -p gpurun_out
for s in 1 2 3 4 8; do
python bench.py --no-cpu-baseline --no-configs --no-stock --steps 5 --warmup 3 --streams $s > gpurun_out/s4_streams$s.log 2> gpurun_out/s4_streams$s.err; echo streams $s rc=$?
python - <<P
import json
for l in open('gpurun_out/s4_streams$s.log'):
    if l.startswith('{'):
        d=json.loads(l); print($s, d['ms_per_frame'], d['e2e']['ms_per_frame'], d.get('eager_ms_per_frame'), d['config']['cuda_graph'])
P
done
