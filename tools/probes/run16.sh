timeout 1500 python -m pytest tests/test_gpu_ranks.py -q -x 2>&1 | tail -40 > gpurun_out/ranks6.log
cat gpurun_out/ranks6.log
