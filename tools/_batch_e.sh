set -u
O=gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "filter_off or baseline_shape or c1_shape" 2>&1 | tail -3) | tee $O/e_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mbp/s step %.3f stream %.3f reduce %.3f' % (d['ms_per_step'], d['step_breakdown_ms']['stream_kernel'], d['step_breakdown_ms']['mixture_and_reduce']))"; }
$B --no-filter 2>/dev/null | tee $O/e_probeall_pipe.json | q probeall_pipe
HYMET_SCREEN_LIB=gpurun_variants/libhs_nopipe.so $B --no-filter 2>/dev/null | q probeall_nopipe
$B --tiny 10000 2>/dev/null | q tiny_main
for v in m2cg m2cgtail; do HYMET_SCREEN_LIB=gpurun_variants/libhs_$v.so $B --tiny 10000 2>/dev/null | q tiny_$v; done
C3="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-extras --workload c2 --sketches 300000 --real 3000 --mbp 1250"
$C3 2>/dev/null | q c3_1250
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file $O/r02_launches_c3.csv $C3 > $O/e_ncu_c3.log 2>&1
tail -2 $O/e_ncu_c3.log | cut -c1-200
