mkdir -p gpurun_out
NP=${1:-4}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NP --steps 10 --warmup 3 > gpurun_out/s14_default_n$NP.log 2> gpurun_out/s14_default_n$NP.err; echo default rc=$?
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/s14_default_n$NP.err | tail -4 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $NP --steps 10 --warmup 3 --pipeline-chunks 1 > gpurun_out/s14_pipe1_n$NP.log 2> gpurun_out/s14_pipe1_n$NP.err; echo pipe1 rc=$?
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29516 bench.py --impl reference --gpus $NP --steps 1 --warmup 0 > gpurun_out/s14_ref_n$NP.log 2> gpurun_out/s14_ref_n$NP.err; echo ref rc=$?
tail -1 gpurun_out/s14_ref_n$NP.log | cut -c1-200
python - <<P
import json
for tag in ('default','pipe1'):
  for l in open('gpurun_out/s14_%s_n$NP.log' % tag):
    if l.startswith('{'):
        d=json.loads(l); print(tag, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['config']['cuda_graph'], d['config'].get('cuda_graph_error'), d['config'].get('gradient_sum','')[:70], d['clocks'])
P
