set -u
O=gpurun_out
S="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-extras --s 10000"
$S > $O/i_s10000.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/i_s10000.json')); print(round(d['value']), d['step_breakdown_ms'], d['reduce'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file $O/r02_launches_s10000.csv $S > $O/i_ncu.log 2>&1
W="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-extras --wta --clusters 250"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file $O/r02_launches_wta.csv $W > $O/i_ncu2.log 2>&1
tail -1 $O/i_ncu.log | cut -c1-100
