set -x
CMD="python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_r2.csv $CMD > gpurun_out/ncu_launch.log 2>&1
CMD2="python tools/probes/c2_once.py 100000000 2"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_msd_partition_bulk|k_checksum|k_msd_count_sort|k_join_bounds|k_join_write|k_build_packed' -s 13 -c 13 -o gpurun_out/prof_c2_r2 $CMD2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench_r2.csv; tail -3 gpurun_out/ncu_full.log
