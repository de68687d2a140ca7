timeout 900 python -m pytest tests/test_gpu_ranks.py -q -x 2>&1 | tail -5
