timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/gputests_r2a.log
cat gpurun_out/gputests_r2a.log
python tools/probes/cfg_prof.py c3 62500000 > gpurun_out/c3_prof_62m_elide.json 2> gpurun_out/c3_prof_62m.err
python tools/probes/cfg_prof.py c4 100000000 > gpurun_out/c4_prof_100m_elide.json 2> gpurun_out/c4_prof_100m.err
python - <<'PY'
import json
for f in ['c3_prof_62m_elide','c4_prof_100m_elide']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['kernel_ms_sum'], dict(list(d['kernels'].items())[:10]))
PY
