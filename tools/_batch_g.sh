set -u
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee $O/g_gputest.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mbp/s step %.3f stream %.3f reduce %.3f' % (d['ms_per_step'], d['step_breakdown_ms']['stream_kernel'], d['step_breakdown_ms']['mixture_and_reduce']))"; }
$B 2>/dev/null | q main_warp_ring
HYMET_SCREEN_LIB=gpurun_variants/libhs_cta_ring.so $B 2>/dev/null | q main_cta_ring
$B 2>/dev/null | q main_warp_ring_again
$B --tiny 10000 2>/dev/null | q tiny_warp_ring
HYMET_SCREEN_LIB=gpurun_variants/libhs_cta_ring.so $B --tiny 10000 2>/dev/null | q tiny_cta_ring
$B --k 31 2>/dev/null | q k31_warp_ring
HYMET_SCREEN_LIB=gpurun_variants/libhs_cta_ring.so $B --k 31 2>/dev/null | q k31_cta_ring
$B --no-filter 2>/dev/null | q probeall_warp_ring
$B --mbp 64 2>/dev/null | q small64_warp_ring
HYMET_SCREEN_LIB=gpurun_variants/libhs_cta_ring.so $B --mbp 64 2>/dev/null | q small64_cta_ring
