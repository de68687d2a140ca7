python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"
python - <<'PY'
import re, json
st={}
for l in open('gpurun_out/bench_n2.err'):
    m=re.search(r'rank (\d+) thread (\w+): query (\d+) (starts|done)(.*)',l)
    if m:
        st[(m.group(1),m.group(2))]=(m.group(3),m.group(4),m.group(5))
print("open jobs:", {k:v for k,v in st.items() if v[1]=='starts'})
try:
    d=json.load(open('gpurun_out/bench_n2.json'))
    print({k:d[k] for k in ['value','ms_per_step','parity','n_gpus']}, d['e2e'])
    print(d['roofline']['kernels_ms_per_step'])
    for k,v in d.get('configs',{}).items(): print(k, {a:v.get(a) for a in ['ms_per_step','value','parity','placement','error','queries_per_s']})
except Exception as e: print("no json", e)
PY
grep -v "\[qce\] rank" gpurun_out/bench_n2.err | grep -v "^\*\|OMP_NUM\|W1018" | tail -12
