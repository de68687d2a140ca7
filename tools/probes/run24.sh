timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_exchange.py -q -x 2>&1 | tail -4
for cfg in "QCE_COUNT_SORT_BULK=1" "QCE_COUNT_SORT_BULK=0"; do
  env $cfg python bench.py --config c2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ck.json 2> gpurun_out/ck.err
  python -c "
import json; d=json.load(open('gpurun_out/ck.json')); k=d['roofline']['kernels_ms_per_step']; print('$cfg', round(d['ms_per_step'],3), 'count_sort', k.get('msd_count_sort'), d['parity']['full_vs_checker'])"
done
