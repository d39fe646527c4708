set -u
O=gpurun_out
# e2e parity of every ingest mode after the hand-off fix (bench checks each mode against the resident result)
(time python bench.py --no-extras --no-cpu > $O/c_bench_e2e.json) 2> $O/c_bench_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c_bench_e2e.json'))
e=d['e2e']; print('value', round(d['value']), 'e2e', round(e['value']), e['ms_per_step'], 'match', e['matches_device_resident'])
for k,v in e['ingest_modes'].items(): print(' ', k, round(v['value']), v['ms_per_step'], v['ok'], v['h2d_bytes_per_step'])
print(' file', {k: d['e2e_file'][k] for k in ('value','ms_per_step','ok','h2d_bytes_per_step')}, d['e2e_file'].get('other_reader_counts'))
PY
HYMET_PACK_STREAMING=1 python bench.py --no-extras --no-cpu > $O/c_bench_e2e_nt.json 2> $O/c_bench_e2e_nt.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c_bench_e2e_nt.json'))
e=d['e2e']; print('NT stores: e2e', round(e['value']), e['ms_per_step'], 'match', e['matches_device_resident'])
for k,v in e['ingest_modes'].items(): print(' ', k, round(v['value']), v['ms_per_step'], v['ok'])
print(' file', d['e2e_file']['value'], d['e2e_file']['ms_per_step'])
PY
(time python tools/cli_bench.py > $O/r02_cli_bench.json) 2> $O/c_cli.err; tail -3 $O/c_cli.err; cat $O/r02_cli_bench.json | cut -c1-1500
(time python tools/stage_bench.py > $O/r02_stage_bench.json) 2> $O/c_stage.err; tail -3 $O/c_stage.err; cat $O/r02_stage_bench.json | cut -c1-1200
(time python tools/f3_cache_bench.py > $O/r02_f3_cache_bench.json) 2> $O/c_f3.err; tail -3 $O/c_f3.err; cat $O/r02_f3_cache_bench.json | cut -c1-1200
