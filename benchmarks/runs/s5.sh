mkdir -p gpurun_out
run() {
tag=$1; shift
python bench.py --no-cpu-baseline --no-configs --no-stock --steps 5 --warmup 3 "$@" > gpurun_out/s5_$tag.log 2> gpurun_out/s5_$tag.err; echo $tag rc=$?
python - <<P
import json
for l in open('gpurun_out/s5_$tag.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', d['ms_per_frame'], d['e2e']['ms_per_frame'], d.get('eager_ms_per_frame'), d['config']['cuda_graph'])
P
}
run p0_s2 --stream-priority 0 --streams 2
run pm1_s2 --stream-priority -1 --streams 2
run pm1_s3 --stream-priority -1 --streams 3
run pm1_s4 --stream-priority -1 --streams 4
run pm1_s2_v4 --stream-priority -1 --streams 2 --kernel-variant 4
run pm5_s3 --stream-priority -5 --streams 3
