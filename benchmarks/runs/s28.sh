mkdir -p gpurun_out
NP=${1:-4}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $NP --steps 10 --warmup 3 > gpurun_out/s28_default_n$NP.log 2> gpurun_out/s28_default_n$NP.err; echo default rc=$?
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/s28_default_n$NP.err | tail -4 | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/s28_default_n$NP.log'):
    if l.startswith('{'):
        d=json.loads(l); print($NP, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['config']['cuda_graph'], d['config'].get('cuda_graph_error'), d['config'].get('gradient_sum','')[:70])
P
