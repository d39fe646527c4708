#!/bin/bash
# One gpurun call: ncu launch list + `--set full` captures of the kernels the roofline keys quote
# (B200_PROFILING.md recipe: plain run first, `--clock-control none`, one GPU).  Outputs -> gpurun_out/.
#   bash tools/gpu_profile.sh <tag> [quick]      quick = launch list + k_stream<21,0> + the Bloom-tier kernel only
set -u
TAG=${1:-r02}
QUICK=${2:-}
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras"
mkdir -p $OUT
$B > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_stream -c 1 \
    -f -o $OUT/${TAG}_stream_full $B > $OUT/${TAG}_ncu_stream.log 2>&1
if [ -z "$QUICK" ]; then
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_probe -c 1 \
    -f -o $OUT/${TAG}_probe_full $B > $OUT/${TAG}_ncu_probe.log 2>&1
fi
$B --tiny 10000 > $OUT/${TAG}_plain_tiny.json 2> $OUT/${TAG}_plain_tiny.err &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_stream -c 1 \
    -f -o $OUT/${TAG}_stream_tiny_full $B --tiny 10000 > $OUT/${TAG}_ncu_stream_tiny.log 2>&1
if [ -z "$QUICK" ]; then
$B --no-filter > $OUT/${TAG}_plain_probeall.json 2> $OUT/${TAG}_plain_probeall.err &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_stream -c 1 \
    -f -o $OUT/${TAG}_stream_probeall_full $B --no-filter > $OUT/${TAG}_ncu_stream_probeall.log 2>&1
fi
ls -la $OUT/${TAG}_*ncu-rep 2>/dev/null
for f in $OUT/${TAG}_ncu_*.log; do echo "$f: $(tail -n 1 $f | cut -c1-160)"; done
