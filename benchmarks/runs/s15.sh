mkdir -p gpurun_out
python benchmarks/configs.py --only c2,c2_graph,c3,c3_graph,c4,c4_graph,c5,c5_graph > gpurun_out/s15_configs.log 2>&1; echo rc=$?
python - <<P
import json
for l in open('gpurun_out/s15_configs.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], d.get('ms_per_frame_fwd_bwd'), d.get('error','')[:300], d.get('peak_mem_gb'))
P
