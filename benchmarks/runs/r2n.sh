mkdir -p gpurun_out
CMD="python benchmarks/variants.py --variants 0 --scene c4 --rounds 1 --iters 2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_fwd_wide' -s 2 -c 1 -f -o gpurun_out/prof_widefwd_r2n $CMD > gpurun_out/r2n_ncu.log 2>&1; echo ncu rc=$?
