#!/bin/bash
# Full ncu captures (with source-level samples) of the front-end kernels that dominate the step. Usage: bash tools/profile_r2c.sh <tag>
set -u
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"adaptive_voxel_kernel|fe_ingest_second_insert|fe_first_filter_insert|fe_emit_tracking|fe_mark_bits|nls_fused_kernel" \
    -s 18 -c 6 -f -o $OUT/${TAG}_front_full python bench.py --steps 2 --warmup 3 --no-extras --pairs 0 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_front_full.ncu-rep
echo done
