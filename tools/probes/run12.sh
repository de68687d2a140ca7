( time python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2> gpurun_out/bench_n1.time
echo "bench n1 rc=$?"; cat gpurun_out/bench_n1.time | tail -4
python tools/probes/cfg_prof.py c4 100000000 > gpurun_out/c4_prof_100m.json 2> gpurun_out/c4_prof_100m.err
python tools/probes/cfg_prof.py c3 62500000 > gpurun_out/c3_prof_62m.json 2> gpurun_out/c3_prof_62m.err
( time timeout 900 python bench.py --config c5 --steps 1 --warmup 0 > gpurun_out/c5_full_n1.json 2> gpurun_out/c5_full_n1.err ) 2> gpurun_out/c5_full_n1.time
echo "c5 full rc=$?"; tail -4 gpurun_out/c5_full_n1.time; tail -3 gpurun_out/c5_full_n1.err
