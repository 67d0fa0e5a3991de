mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s12_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/s12_smoke.log | cut -c1-300
python -m pytest tests -m gpu -q -x > gpurun_out/s12_pytest.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/s12_pytest.log | cut -c1-300
python bench.py > gpurun_out/s12_bench.log 2> gpurun_out/s12_bench.err; echo bench rc=$?
tail -3 gpurun_out/s12_bench.err | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/s12_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('ms_per_frame','eager_ms_per_frame','without_stale_tail_ms_per_frame','stock_api_ms_per_frame','single_stream_ms_per_frame')}, d['e2e']['ms_per_frame'], d['config']['cuda_graph'], d['config']['cuda_graph_error'])
        print({k:(v['ms_per_frame']) for k,v in d['configs'].items() if isinstance(v,dict)})
P
