timeout 1500 python -m pytest tests/test_gpu_ranks.py -q 2>&1 | tail -60 > gpurun_out/ranks4.log
QCE_TRACE=1 python bench.py --config c2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c2_trace.json 2> gpurun_out/c2_trace.err
QCE_TRACE=1 python tools/probes/cfg_prof.py c3 8000000 > gpurun_out/c3_prof_8m_b.json 2> gpurun_out/c3_trace.err
cat gpurun_out/ranks4.log; grep -c "arena" gpurun_out/c2_trace.err; grep "arena" gpurun_out/c2_trace.err | tail -12; grep "batch of" gpurun_out/c2_trace.err | tail -12; grep "arena" gpurun_out/c3_trace.err | tail; grep "batch of" gpurun_out/c3_trace.err | tail -8
