set -u
O=gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py -x -q 2>&1 | tail -4) > $O/a_tests.log
cat $O/a_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mbp/s step %.3f stream %.3f' % (d['ms_per_step'], d['step_breakdown_ms']['stream_kernel']))"; }
$B 2>/dev/null | tee $O/a_main.json | q main
for v in notail nosleep; do HYMET_SCREEN_LIB=gpurun_variants/libhs_$v.so $B 2>/dev/null | q $v; done
$B --tiny 10000 2>/dev/null | q tiny_main
$B --no-filter 2>/dev/null | q probeall_main
HYMET_SCREEN_LIB=gpurun_variants/libhs_nopf.so $B --no-filter 2>/dev/null | q probeall_nopf
$B --k 31 2>/dev/null | q k31_main
python tools/e2e_debug.py 1000 2>&1 | tee $O/a_e2e_debug.log
