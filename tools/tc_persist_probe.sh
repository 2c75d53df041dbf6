for p in 0 1; do
  echo "persist=$p"
  SGPU_FIR_TC_PERSIST=$p timeout 100 python tools/tc_probe.py 27 2>&1 | grep -v "over the whole" | grep "512.*bf16-c2"
  SGPU_FIR_TC_PERSIST=$p ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fir_tc -c 2 --csv --log-file gpurun_out/r1w_traffic_p$p.csv python bench.py --workload fir --steps 1 --warmup 1 --no-cpu --no-e2e --no-check > /dev/null 2>&1
  tail -3 gpurun_out/r1w_traffic_p$p.csv | cut -d, -f13-
done
