mkdir -p gpurun_out
CMD="python benchmarks/variants.py --variants 0 --scene c4 --rounds 1 --iters 2"
$CMD > gpurun_out/r2f_plain.log 2>&1; echo plain rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_bwd_wide|raster_fwd_fast' -s 2 -c 2 -f -o gpurun_out/prof_wide_r2f $CMD > gpurun_out/r2f_ncu.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/prof_wide_r2f*
