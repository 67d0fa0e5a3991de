mkdir -p gpurun_out
summ='import sys, json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ("n_gpus","ms_per_step","ms_per_frame","scaling")}, "e2e", d["e2e"]["ms_per_step"], d["config"]["views_per_rank"])'
for mode in "" "--no-reduce-early"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 3 $mode > gpurun_out/r2h_n8$mode.log 2>&1; echo rc=$?
tail -1 gpurun_out/r2h_n8$mode.log | python -c "$summ"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 4 --warmup 3 --workload c5 --total-views 64 > gpurun_out/r2h_c5_n8.log 2>&1; echo rc=$?
tail -1 gpurun_out/r2h_c5_n8.log | python -c "$summ"
python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k "8" 2>&1 | tail -6
