mkdir -p gpurun_out
NP=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 benchmarks/allreduce_probe.py > gpurun_out/s7_probe_n$NP.log 2> gpurun_out/s7_probe_n$NP.err; echo probe rc=$?
tail -5 gpurun_out/s7_probe_n$NP.err | cut -c1-400
tail -2 gpurun_out/s7_probe_n$NP.log | cut -c1-1500
