# Evidence capture for profiles/ (run under gpurun, one GPU). Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
# 1. memory checker on the kernels added late in the round (loop closure, streaming front end)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fcsm.py tests/test_gpu_streaming.py -x -q -m gpu > gpurun_out/r1_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r1_memcheck.log
# 2. every launch of the default bench command with its device time and DRAM bytes
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r1b_launches_raw.csv python bench.py --steps 2 --warmup 3 --cpu-sample 8 > gpurun_out/r1b_launch.log 2>&1
# 3. full counters of the loop-closure search kernel (one launch = 256 pairs)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fcsm_search -s 1 -c 1 -o gpurun_out/r1b_fcsm \
    python tools/bench_loop_closure.py --steps 1 --warmup 1 --cpu-pairs 1 > gpurun_out/r1b_fcsm.log 2>&1
ls -la gpurun_out | tail -8
tail -3 gpurun_out/r1_memcheck.log
