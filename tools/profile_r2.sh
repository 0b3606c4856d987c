#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): launch list with DRAM traffic of one C4 iteration, full ncu captures of
# the tile GEMM on level-0 (small fronts) and top-of-tree launches and of the assembly kernel.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
# every launch of one iteration with device time and DRAM bytes (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_ncu_launches_c4.csv python tools/one_iter.py C4 64 > gpurun_out/one_iter_ncu.log 2>&1
python tools/ncu_traffic_summary.py gpurun_out/r2_ncu_launches_c4.csv 1 gpurun_out/r2_ncu_dram_traffic_c4.json > gpurun_out/r2_ncu_dram_traffic_c4.txt 2>&1
# full captures: GEMM launches 0-3 are level 0 (pivot panels / left updates of ~16k leaf fronts), 6-9 the block-doubling
# inverse and the level-0 Schur scatter; launches from ~380 on are the top of the tree
ncu --set full --clock-control none --import-source on -k regex:gemm_tile -s 0 -c 10 -o gpurun_out/r2_gemm_level0 \
    python tools/one_iter.py C4 64 > gpurun_out/one_iter_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tile -s 392 -c 3 -o gpurun_out/r2_gemm_top \
    python tools/one_iter.py C4 64 > gpurun_out/one_iter_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"assemble_g|station_sum|init_normals" -c 3 -o gpurun_out/r2_assemble \
    python tools/one_iter.py C4 64 > gpurun_out/one_iter_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
