mkdir -p gpurun_out
python benchmarks/profile_config.py c3 --top 40 > gpurun_out/s27_c3_profile.log 2>&1; echo rc=$?
grep -v "Warning\|_warn_once" gpurun_out/s27_c3_profile.log | tail -41 | cut -c1-150
