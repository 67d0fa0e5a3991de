set -x
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-configs --no-stock --steps 3 --warmup 3 > gpurun_out/s3_bench.log 2> gpurun_out/s3_bench.err; echo bench rc=$?
grep -v "^$" gpurun_out/s3_bench.err | tail -40 | cut -c1-300
tail -1 gpurun_out/s3_bench.log | cut -c1-300
