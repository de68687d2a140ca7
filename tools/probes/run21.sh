( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err ) 2> gpurun_out/bench_n4.time
echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n4.time
python - <<'PY'
import json
s=open('gpurun_out/bench_n4.json').read().strip()
try:
    d=json.loads(s.splitlines()[-1])
    print({k:d.get(k) for k in ['value','ms_per_step','parity','n_gpus']}, d['e2e']['ms_per_step'])
    print(d['roofline']['kernels_ms_per_step'])
    for k,v in d.get('configs',{}).items(): print(k, {a:v.get(a) for a in ['ms_per_step','value','parity','placement','error','queries_per_s','load_s']})
except Exception as e: print("no json", e)
PY
grep -v "^\*\|OMP_NUM\|W1018" gpurun_out/bench_n4.err | tail -8
