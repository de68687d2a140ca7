timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2> gpurun_out/bench_n1.time
echo "bench n1 rc=$?"; tail -3 gpurun_out/bench_n1.time
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','parity']}, d['e2e']['ms_per_step'])
print(d['roofline']['kernels_ms_per_step']); print(d['roofline']['kernels_frac_of_peak'], d['roofline']['stage'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])
for k,v in d.get('configs',{}).items(): print(k, {a:v.get(a) for a in ['ms_per_step','value','parity','error','queries_per_s']})
PY
