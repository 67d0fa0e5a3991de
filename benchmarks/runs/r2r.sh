mkdir -p gpurun_out
python bench.py > gpurun_out/r2r_bench.log 2>&1; echo bench rc=$?
tail -1 gpurun_out/r2r_bench.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_ref.log 2>&1; echo ref rc=$?
tail -1 gpurun_out/r2r_ref.log | cut -c1-300
python benchmarks/timeline.py > gpurun_out/r2r_timeline.log 2>&1; echo timeline rc=$?
tail -25 gpurun_out/r2r_timeline.log | cut -c1-200
