#!/bin/bash
# Launch list of the bench step (serialised by ncu: per-kernel time + DRAM bytes), then a clean bench line.
# Usage under gpurun: bash tools/profile_r2b.sh <tag>
set -u
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 400 -c 300 --csv \
    --log-file $OUT/${TAG}_launches_dram_raw.csv python bench.py --steps 4 --warmup 3 --no-extras > $OUT/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py $OUT/${TAG}_launches_dram_raw.csv > $OUT/${TAG}_launch_summary.json 2>> $OUT/${TAG}_ncu.log
python bench.py --steps 50 --no-extras > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo done
