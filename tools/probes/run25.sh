timeout 900 python -m pytest tests/test_gpu_primitives.py -q -x 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_queries.py -q -x -k "zipf or heavy or elided or c4" 2>&1 | tail -4
python tools/probes/cfg_prof.py c4 100000000 > gpurun_out/c4_prof_100m_big.json 2> gpurun_out/c4_prof_100m.err
QCE_MSD_BIG=0 python tools/probes/cfg_prof.py c4 100000000 > gpurun_out/c4_prof_100m_nobig.json 2>> gpurun_out/c4_prof_100m.err
python - <<'PY'
import json
for f in ['c4_prof_100m_big','c4_prof_100m_nobig']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['ms_per_step'],3), round(d['kernel_ms_sum'],3), dict(list(d['kernels'].items())[:12]))
PY
tail -3 gpurun_out/c4_prof_100m.err
