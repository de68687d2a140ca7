QCE_MSD_BULK=0 timeout 1500 python -m pytest tests/test_gpu_ranks.py -q 2>&1 | tail -80 > gpurun_out/ranks3.log
timeout 600 python -m pytest tests/test_gpu_primitives.py -q -k "sort or msd" 2>&1 | tail -15 > gpurun_out/bulk_sort_tests.log
python tools/probes/cfg_prof.py c3 8000000 > gpurun_out/c3_prof_8m.json 2> gpurun_out/c3_prof.err
python tools/probes/cfg_prof.py c4 20000000 > gpurun_out/c4_prof_20m.json 2>> gpurun_out/c3_prof.err
QCE_MSD_BULK=0 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_plain.json 2> gpurun_out/c2_plain.err
QCE_MSD_BULK=1 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bulk.json 2> gpurun_out/c2_bulk.err
cat gpurun_out/ranks3.log gpurun_out/bulk_sort_tests.log; tail -5 gpurun_out/c3_prof.err gpurun_out/c2_bulk.err
