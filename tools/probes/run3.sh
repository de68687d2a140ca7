timeout 1200 python -m pytest tests/test_gpu_ranks.py -x -q 2>&1 | tail -60 > gpurun_out/ranks2.log
timeout 600 python bench.py --steps 3 --warmup 1 --rows 20000000 --c3-rows 8000000 --c5-scale 0.02 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "bench rc=$?" >> gpurun_out/ranks2.log
tail -c 3000 gpurun_out/bench_small.err >> gpurun_out/ranks2.log
cat gpurun_out/ranks2.log
