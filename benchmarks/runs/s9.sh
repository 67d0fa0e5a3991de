mkdir -p gpurun_out
C3="python benchmarks/variants.py --variants 0 --scene c3 --rounds 1 --iters 2"
$C3 > gpurun_out/s9_c3_plain.log 2>&1; echo c3 plain rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'raster_fwd_fast|raster_bwd_fast' -s 4 -c 2 -f -o gpurun_out/prof_c3_r02b $C3 > gpurun_out/s9_ncu_c3.log 2>&1; echo c3 capture rc=$?
ls -la gpurun_out/prof_c3_r02b*
