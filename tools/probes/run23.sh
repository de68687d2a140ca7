for st in 1 4 8; do
  QCE_STREAMS=$st python bench.py --config c5 --c5-scale 0.0005 --steps 5 --warmup 2 > gpurun_out/c5_tiny_$st.json 2> gpurun_out/c5_tiny_$st.err
  python -c "
import json; d=json.loads(open('gpurun_out/c5_tiny_$st.json').read().strip().splitlines()[-1]); print('streams $st', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), d['parity'], d['detail']['relation_rows'][0], d['detail']['relation_rows'][-1])"
done
CMD2="python tools/probes/c2_once.py 100000000 2"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_checksum|k_msd_count_sort|k_msd_partition' -s 10 -c 10 -o gpurun_out/prof_c2_r2b $CMD2 > gpurun_out/ncu_full_b.log 2>&1
ls -la gpurun_out/prof_c2_r2b.ncu-rep
