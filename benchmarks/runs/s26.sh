mkdir -p gpurun_out
python benchmarks/profile_config.py c4 --top 30 > gpurun_out/s26_c4_profile.log 2>&1; echo rc=$?
grep -v Warning gpurun_out/s26_c4_profile.log | tail -32 | cut -c1-170
