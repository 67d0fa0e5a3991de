mkdir -p gpurun_out
run() {
tag=$1; shift
python bench.py --no-cpu-baseline --no-configs --no-stock --steps 6 --warmup 3 "$@" > gpurun_out/s18_$tag.log 2> gpurun_out/s18_$tag.err; echo $tag rc=$?
tail -2 gpurun_out/s18_$tag.err | cut -c1-200
python - <<P
import json
for l in open('gpurun_out/s18_$tag.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['ms_per_frame'],4), round(d['e2e']['ms_per_frame'],4), round(d.get('eager_ms_per_frame') or 0,4), round(d['single_stream_ms_per_frame'],4), d['stage_ms_per_frame']['gs_raster_bwd'], d['stage_ms_per_frame']['gs_raster_fwd'])
P
}
run base
run b28 --kernel-variant 32
run b24 --kernel-variant 64
run b20 --kernel-variant 96
run b24f7 --kernel-variant 192
run b24f6 --kernel-variant 320
run b28f7 --kernel-variant 160
