python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/bench_n2.err
timeout 900 python -m pytest tests/test_gpu_ranks.py -q -k "scaled or reference_main" 2>&1 | tail -5
