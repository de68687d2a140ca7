timeout 1500 python -m pytest tests/test_gpu_ranks.py -q 2>&1 | tail -70 > gpurun_out/ranks5.log
cat gpurun_out/ranks5.log
