set -u
O=gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "file_streaming or filter or bloom or baseline_shape" 2>&1 | tail -4) > $O/b_tests.log
cat $O/b_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mbp/s step %.3f stream %.3f' % (d['ms_per_step'], d['step_breakdown_ms']['stream_kernel']))"; }
$B --tiny 10000 2>/dev/null | q tiny_main_gate_notail
for v in m2tail m2nogate m2tailnogate; do HYMET_SCREEN_LIB=gpurun_variants/libhs_$v.so $B --tiny 10000 2>/dev/null | q tiny_$v; done
$B --no-filter 2>/dev/null | q probeall_main
python tools/e2e_debug.py 1000 2>&1 | tee $O/b_e2e_debug.log
