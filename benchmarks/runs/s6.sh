mkdir -p gpurun_out
run() {
tag=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 "$@" > gpurun_out/s6_$tag.log 2> gpurun_out/s6_$tag.err; echo $tag rc=$?
tail -3 gpurun_out/s6_$tag.err | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/s6_$tag.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['cuda_graph'], d['config'].get('cuda_graph_error'), d['step_tail_ms'])
P
}
run graph
run nograph --no-graph
run nograph_noearly --no-graph --no-reduce-early
