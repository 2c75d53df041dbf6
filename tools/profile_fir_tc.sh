#!/bin/bash
# Runs on the GPU box (gpurun): evidence for the tensor-core FIR -- launch list of the default bench command,
# one `ncu --set full` capture of the kernel at 2^26 samples, DRAM traffic of one launch at the bench's full size.
# The plain commands run first (a number printed under ncu is never a bench value).  usage: tools/profile_fir_tc.sh [tag]
set -u
tag=${1:-r1t}
out=gpurun_out
mkdir -p $out
python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_bench_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_fir.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_bench_ncu.log 2>&1
export EXPLORE_REPS=1 EXPLORE_WARM=1
python tools/explore.py fir:512,26 > $out/${tag}_explore_plain.log 2>&1 || { echo "plain explore failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k 'regex:fir_' -o $out/prof_${tag}_full -f \
    python tools/explore.py fir:512,26 > $out/${tag}_explore_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:fir_' -c 3 --csv \
    --log-file $out/${tag}_traffic_fir.csv python bench.py --workload fir --steps 1 --warmup 3 --no-cpu --no-e2e --no-check \
    > $out/${tag}_traffic_fir.log 2>&1
tail -2 $out/${tag}_explore_plain.log; tail -3 $out/${tag}_traffic_fir.csv | cut -d, -f13-
