mkdir -p gpurun_out
summ='import sys, json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ("n_gpus","ms_per_step","ms_per_frame","scaling")}, "e2e", d["e2e"]["ms_per_step"], d["step_tail_ms"])'
for mode in "" "--no-reduce-early"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 $mode > gpurun_out/r2i_n8$mode.log 2>&1; echo rc=$?
tail -1 gpurun_out/r2i_n8$mode.log | python -c "$summ"
done
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-stock > gpurun_out/r2i_n1.log 2>&1
tail -1 gpurun_out/r2i_n1.log | python -c "$summ"
