mkdir -p gpurun_out
python -m pytest tests/test_gpu_renderer.py tests/test_gpu_static_render.py tests/test_gpu_parity_at_size.py -m gpu -q -x > gpurun_out/s25_pytest.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/s25_pytest.log | cut -c1-300
python benchmarks/configs.py --only c4,c4_graph > gpurun_out/s25_c4.log 2>&1; echo rc=$?
python - <<P
import json
for l in open('gpurun_out/s25_c4.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], d.get('ms_per_frame_fwd_bwd'), d.get('error','')[:200], d.get('peak_mem_gb'))
P
mkdir -p gpurun_out
python benchmarks/profile_config.py c4 --top 30 > gpurun_out/s26_c4_profile.log 2>&1; echo rc=$?
grep -v Warning gpurun_out/s26_c4_profile.log | tail -32 | cut -c1-170
