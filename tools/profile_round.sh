#!/bin/bash
# Runs on the GPU box (gpurun): launch list of the default bench, one `ncu --set full` capture of
# every hot kernel at a reduced size, and DRAM traffic of one launch of every bench workload at
# full size.  Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.md.
# usage: tools/profile_round.sh <tag>
set -u
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
K='regex:fir_|iir_|carry|walk|autocorr_kernel'
if [ "${2:-all}" != "traffic" ]; then
# 1. the bench command must exit 0 without ncu first
python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_bench_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_fir.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_bench_ncu.log 2>&1
# 2. full counter sets, reduced sizes (ncu replays every kernel ~40 times)
export EXPLORE_REPS=1 EXPLORE_WARM=1
python tools/explore.py fir:512,26 decim:256,8,256,20 interp:128,4,256,20 iir:65536,12 iirscan:27 autocorr:64,16,256,20 > $out/${tag}_explore_plain.log 2>&1 || { echo "plain explore failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k "$K" -o $out/prof_${tag}_full -f \
    python tools/explore.py fir:512,26 decim:256,8,256,20 interp:128,4,256,20 iir:65536,12 iirscan:27 autocorr:64,16,256,20 > $out/${tag}_explore_ncu.log 2>&1
fi
# 3. DRAM traffic per launch at the bench's full sizes (one launch each)
for w in fir decim interp iir_batch iir_scan autocorr; do
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" -c 3 --csv \
        --log-file $out/${tag}_traffic_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu --no-e2e --no-check \
        > $out/${tag}_traffic_$w.log 2>&1
done
ls -la $out | tail -20
