set -u
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee $O/r02_gputest_final.log
(time python bench.py > $O/r02_bench_n1_final.json) 2> $O/f_bench.err; tail -2 $O/f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_final.json'))
print('value', round(d['value']), d['ms_per_step'], d['step_breakdown_ms'])
print('roofline', d['roofline']['bound'], d['roofline']['frac'], d['roofline']['hbm']['frac'])
e=d['e2e']; print('e2e', round(e['value']), e['ms_per_step'], e['matches_device_resident'], {k:(round(v['value']),v['ok']) for k,v in e['ingest_modes'].items()})
print('file', round(d['e2e_file']['value']), 'packed', round(d['e2e_packed']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['parity_on_sample'])
print('c3', round(d['c3_single_gpu']['value']), 'tiny', round(d['tiny_db']['value']), 'probe', d['probe_kernel']['probes_per_s'], d['probe_kernel']['frac'])
PY
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mbp/s step %.3f stream %.3f reduce %.3f' % (d['ms_per_step'], d['step_breakdown_ms']['stream_kernel'], d['step_breakdown_ms']['mixture_and_reduce']))"; }
$B --no-filter 2>/dev/null | q probeall
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2>/dev/null; cut -c1-400 $O/r02_bench_reference.json
