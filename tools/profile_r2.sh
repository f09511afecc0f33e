#!/bin/bash
# Round-2 profile collection (run under gpurun from the repo root): launch list with per-kernel DRAM bytes, full captures of the
# dominant kernels, the correlative kernel on the configs[2] shape. Outputs under gpurun_out/; summaries are copied to profiles/.
set -u
OUT=gpurun_out
mkdir -p $OUT
# 1. serialised launch list with duration + DRAM bytes per kernel (a few passes per kernel)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 500 -c 260 --csv \
    --log-file $OUT/r2_launches_dram.csv python bench.py --steps 4 --warmup 3 --no-extras > $OUT/r2_ncu_a.log 2>&1
# 2. full captures of the front-end kernels that dominate the step
ncu --set full --clock-control none --import-source on -k regex:"fe_ingest_second_insert|fe_first_filter_insert|adaptive_voxel_kernel|nls_fused_kernel" \
    -s 16 -c 4 -o $OUT/r2_front_full python bench.py --steps 2 --warmup 3 --no-extras > $OUT/r2_ncu_b.log 2>&1
# 3. the correlative kernel (configs[2] shape: 0.05 m grid, 0.15 m / 1 deg window) alone
ncu --set full --clock-control none --import-source on -k regex:"rtcsm_score_kernel" -c 1 -o $OUT/r2_rtcsm_full \
    python tools/profile_rtcsm.py > $OUT/r2_ncu_c.log 2>&1
python tools/profile_rtcsm.py > $OUT/r2_rtcsm_timing.json 2>> $OUT/r2_ncu_c.log
echo done
