#!/bin/bash
# Round-2 evidence run, second half (64 x 64 tile shape, rewritten gather / pivot-tile kernels) on one B200 under gpurun:
# launch list with DRAM traffic of one C4 iteration; full ncu captures of the 64 x 64 tile GEMM on a level-0 Schur launch,
# a mid-tree Z21 launch, and of the gather kernel.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2b_ncu_launches_c4.csv python tools/one_iter.py C4 64 > gpurun_out/one_iter_ncu.log 2>&1
python tools/ncu_traffic_summary.py gpurun_out/r2b_ncu_launches_c4.csv 1 gpurun_out/r2b_ncu_dram_traffic_c4.json > gpurun_out/r2b_ncu_dram_traffic_c4.txt 2>&1
# GEMM launches in plan order (profiles/r2_tile_shapes_c4.txt): 5 = level-0 Schur scatter, 6 = level-1 pivot panel, 7 = level-1
# Schur scatter (all on 64 x 64 tiles, 3 CTAs/SM); 426 = Z21 of tree level 6 (64 x 64 tiles, K ~ 1300); gather launch 8 = level 5
ncu --set full --clock-control none --import-source on -k regex:gemm_tile -s 5 -c 3 -o gpurun_out/r2b_gemm64_level0 \
    python tools/one_iter.py C4 64 > gpurun_out/one_iter_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tile -s 426 -c 1 -o gpurun_out/r2b_gemm64_mid \
    python tools/one_iter.py C4 64 > gpurun_out/one_iter_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 8 -c 1 -o gpurun_out/r2b_gather \
    python tools/one_iter.py C4 64 > gpurun_out/one_iter_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
