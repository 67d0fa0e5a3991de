mkdir -p gpurun_out
NP=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29513 tests/multi_gpu_worker.py > gpurun_out/s8_worker_n$NP.log 2>&1; echo worker rc=$?
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/s8_worker_n$NP.log | tail -12 | cut -c1-300
run() {
tag=$1; shift
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NP --steps 5 --warmup 3 "$@" > gpurun_out/s8_${tag}_n$NP.log 2> gpurun_out/s8_${tag}_n$NP.err; echo $tag rc=$?
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/s8_${tag}_n$NP.err | tail -6 | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/s8_${tag}_n$NP.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['cuda_graph'], d['config'].get('cuda_graph_error'), d['config'].get('gradient_sum'))
P
}
run symm
run symm_noearly --no-reduce-early
run nccl --nccl-bucket
