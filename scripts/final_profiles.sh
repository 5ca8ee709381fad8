#!/bin/bash
# One-GPU evidence run for profiles/ (run on the GPU box through gpurun; everything lands in gpurun_out/r2_*).
KR='khop_|tree_rows|rows_|expand_level|roots_assign|lid_clear|level_snapshot|batch_gather|linear_tf32|halo|stage_claim'
LIGHT="--steps 1 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline --no-full-graph"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_reference_arm_1gpu.json 2> gpurun_out/r2_reference_arm_1gpu.err
python bench.py --steps 20 --warmup 5 --streams 1 --no-cpu-baseline --no-full-graph > gpurun_out/r2_bench_1gpu_streams1.json 2> /dev/null
python bench.py --steps 20 --warmup 5 --streams 2 --no-cpu-baseline --no-full-graph > gpurun_out/r2_bench_1gpu_streams2.json 2> /dev/null
# DRAM bytes + time of every launch of one step (roofline.traffic)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name regex:"$KR" -c 400 --csv \
    --log-file gpurun_out/r2_traffic.csv python bench.py $LIGHT > /dev/null 2> gpurun_out/r2_traffic.err
# launch list of the same command (shares)
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"$KR" -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py $LIGHT > /dev/null 2> gpurun_out/r2_launches.err
# full metric set (+ source) of the hot kernels of the first two steps
ncu --set full --import-source on --clock-control none \
    --kernel-name regex:"khop_tile_kernel|batch_gather_async_kernel|linear_tf32x3_kernel|rows_sort_kernel|tree_rows_kernel|expand_level_kernel" \
    -c 26 -o gpurun_out/r2_full python bench.py $LIGHT > /dev/null 2> gpurun_out/r2_full.err
python bench.py --workload mag-like --steps 10 --warmup 3 --no-cpu-baseline --no-full-graph > gpurun_out/r2_mag_like_1gpu.json 2> gpurun_out/r2_mag_like_1gpu.err
python bench.py --workload g2b-small --steps 5 --warmup 2 > gpurun_out/r2_g2b_small_1gpu.json 2> gpurun_out/r2_g2b_small_1gpu.err
echo done
