set -u
out=gpurun_out
timeout 600 python -m pytest tests/test_fir_tc_gpu.py tests/test_fir_gpu.py "tests/test_fullsize_gpu.py::test_config2_fir_full_stream" -x -q -m gpu > $out/t_tc2.log 2>&1; tail -5 $out/t_tc2.log
timeout 300 python bench.py --steps 10 --warmup 3 > $out/r1t_bench.json 2> $out/r1t_bench.err || tail -5 $out/r1t_bench.err
head -c 1500 $out/r1t_bench.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r1t_launches_bench_fir.csv python bench.py --steps 2 --warmup 3 --no-cpu > $out/r1t_bench_ncu.log 2>&1
export EXPLORE_REPS=1 EXPLORE_WARM=1
python tools/explore.py fir:512,26 > $out/r1t_explore_plain.log 2>&1 || echo "plain explore failed"
ncu --set full --clock-control none --import-source on -k 'regex:fir_' -o $out/prof_r1t_full -f python tools/explore.py fir:512,26 > $out/r1t_explore_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:fir_' -c 3 --csv --log-file $out/r1t_traffic_fir.csv python bench.py --workload fir --steps 1 --warmup 3 --no-cpu --no-e2e --no-check > $out/r1t_traffic_fir.log 2>&1
tail -3 $out/r1t_explore_plain.log; ls -la $out | tail -8
