mkdir -p gpurun_out
NP=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --workload c5 --total-views 64 --steps 6 --warmup 3 > gpurun_out/s21_c5_n$NP.log 2> gpurun_out/s21_c5_n$NP.err; echo c5 rc=$?
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/s21_c5_n$NP.err | tail -3 | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/s21_c5_n$NP.log'):
    if l.startswith('{'):
        d=json.loads(l); print($NP, round(d['ms_per_step'],3), round(d['ms_per_frame'],4), round(d['e2e']['ms_per_step'],3), d['scaling'], d['config']['cuda_graph'], d['config']['views_per_rank'], d['value'])
P
