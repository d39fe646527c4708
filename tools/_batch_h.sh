set -u
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee $O/r02_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time python bench.py > $O/r02_bench_n1_final.json) 2> $O/h_bench.err; tail -2 $O/h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_final.json'))
print('value', round(d['value']), d['ms_per_step'], d['step_breakdown_ms'])
print('roofline', d['roofline']['bound'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['ncu_static'])
e=d['e2e']; print('e2e', round(e['value']), e['ms_per_step'], e['matches_device_resident'], {k:(round(v['value']),v['ok']) for k,v in e['ingest_modes'].items()})
print('file', round(d['e2e_file']['value']), 'packed', round(d['e2e_packed']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['parity_on_sample'])
print('c3', round(d['c3_single_gpu']['value']), 'tiny', round(d['tiny_db']['value']), 'probe', d['probe_kernel']['probes_per_s'], d['probe_kernel']['frac'], d['gpu_launches'])
PY
