mkdir -p gpurun_out
summ='import sys, json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ("n_gpus","ms_per_step")}, "e2e", d["e2e"]["ms_per_step"], d["step_tail_ms"])'
for mode in "--background-ctas 0" "--background-ctas 4" "--background-ctas 8" "--background-ctas 16" "--no-reduce-early"; do
echo "== $mode"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 $mode > gpurun_out/r2j_n2.log 2>&1; echo rc=$?
tail -1 gpurun_out/r2j_n2.log | python -c "$summ" || tail -5 gpurun_out/r2j_n2.log
done
