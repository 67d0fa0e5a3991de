mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -12
for mode in "" "--no-reduce-early"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 $mode > gpurun_out/r2g_n2$mode.log 2>&1; echo rc=$?
tail -1 gpurun_out/r2g_n2$mode.log | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('n_gpus','ms_per_step','ms_per_frame')}, d['e2e']['ms_per_step'], d['config']['parallelism'][:80])"
done
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-stock > gpurun_out/r2g_n1.log 2>&1
tail -1 gpurun_out/r2g_n1.log | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('n_gpus','ms_per_step','ms_per_frame')}, d['e2e']['ms_per_step'])"
