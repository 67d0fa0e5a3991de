mkdir -p gpurun_out
python bench.py --workload c5 --total-views 24 --steps 3 --no-cpu-baseline --no-configs --no-stock > gpurun_out/s20_c5.log 2> gpurun_out/s20_c5.err; echo rc=$?
tail -3 gpurun_out/s20_c5.err | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/s20_c5.log'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],3), round(d['ms_per_frame'],4), round(d['e2e']['ms_per_frame'],4), d['scaling'], d['config']['cuda_graph'], d['config']['cuda_graph_error'], d['config']['views_per_rank'], d['config']['workload'][:80])
P
