mkdir -p gpurun_out
python benchmarks/configs.py --only c1,c1_graph,c2,c2_graph,c3,c3_graph,c4,c4_graph,c5,c5_graph > gpurun_out/configs_r02c.log 2>&1; echo configs rc=$?
C4="python benchmarks/variants.py --variants 0 --scene c4 --rounds 1 --iters 2"
$C4 > gpurun_out/c4_plain_r02c.log 2>&1; echo c4 plain rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'raster_bwd_wide|raster_fwd_fast' -s 4 -c 2 -f -o gpurun_out/prof_c4_r02c $C4 > gpurun_out/ncu_c4_r02c.log 2>&1; echo c4 capture rc=$?
ls -la gpurun_out/prof_c4_r02c*
