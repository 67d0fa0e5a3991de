mkdir -p gpurun_out
NP=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29513 tests/multi_gpu_worker.py > gpurun_out/s13_worker_n$NP.log 2>&1; echo worker rc=$?
grep "rel l2" gpurun_out/s13_worker_n$NP.log | cut -c1-200
run() {
tag=$1; shift
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NP --steps 8 --warmup 3 "$@" > gpurun_out/s13_${tag}_n$NP.log 2> gpurun_out/s13_${tag}_n$NP.err; echo $tag rc=$?
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/s13_${tag}_n$NP.err | tail -4 | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/s13_${tag}_n$NP.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['config']['cuda_graph'], d['config'].get('cuda_graph_error'), d['config'].get('gradient_sum','')[:60])
P
}
run pipe4 --symmetric-from 2 --pipeline-chunks 4
run pipe1 --symmetric-from 2 --pipeline-chunks 1
run pipe8 --symmetric-from 2 --pipeline-chunks 8
run nccl --nccl-bucket
