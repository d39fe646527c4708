set -u
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee $O/d_gputest.log
(timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_parity.py 2>&1 | grep -E "parity|rror|rank") | tee $O/r02_dist_parity_n2.log
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/d_bench_n2.json) 2> $O/d_bench_n2.err; tail -3 $O/d_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d_bench_n2.json'))
print('value', round(d['value']), d['ms_per_step'], d['step_breakdown_ms']); print(d['scaling_loss']); print(d['parity_vs_single']); print(d['c2_weak']['ms_per_step'])
PY
HYMET_SCREEN_SYNC_FREE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras > $O/d_bench_n2_sync.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/d_bench_n2_sync.json'))
print('SYNC (old) value', round(d['value']), d['ms_per_step'], d['step_breakdown_ms']); print(d['scaling_loss']['efficiency_vs_same_run_single_gpu'])
PY
