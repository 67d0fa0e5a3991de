mkdir -p gpurun_out
run() {
tag=$1; shift
python bench.py --no-cpu-baseline --no-configs --no-stock --steps 6 --warmup 3 "$@" > gpurun_out/s22_$tag.log 2> gpurun_out/s22_$tag.err; echo $tag rc=$?
tail -2 gpurun_out/s22_$tag.err | cut -c1-200
python - <<P
import json
for l in open('gpurun_out/s22_$tag.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['ms_per_frame'],4), round(d['e2e']['ms_per_frame'],4), round(d.get('eager_ms_per_frame') or 0,4), d['config']['cuda_graph'])
P
}
run base
run stagger --stagger
run stagger3 --stagger --streams 3
