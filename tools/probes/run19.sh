( time python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2> gpurun_out/bench_n1.time
echo "bench n1 rc=$?"; tail -3 gpurun_out/bench_n1.time
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_arm.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','parity']}, d['e2e']['ms_per_step'], d['same_workload_as_reference_arm'])
print(d['roofline']['kernels_ms_per_step']); print(d['roofline']['kernels_frac_of_peak'], d['roofline']['stage'], d['roofline']['kernel'], d['roofline']['frac'])
for k,v in d.get('configs',{}).items(): print(k, {a:v.get(a) for a in ['ms_per_step','value','parity','placement','error','queries_per_s','load_s','sorted_run_cache']})
PY
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
